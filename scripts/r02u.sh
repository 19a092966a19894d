set -x
tag=$1
for w in c5 c4 c5solved; do
timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu --extras none > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; tail -1 gpurun_out/${tag}_bench_$w.err
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
