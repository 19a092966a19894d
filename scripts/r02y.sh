set -x
tag=$1
timeout 600 python -m pytest tests/test_gpu_qform.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for w in c4 c5solved; do
timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --extras none > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; tail -1 gpurun_out/${tag}_bench_${w}.err
done
