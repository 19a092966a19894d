set -x
tag=$1
timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -5
for c in 1 2 4; do
DLB_OZAKI_CLUSTER=$c timeout 300 python bench.py --workload c4wide_i8 --steps 10 --warmup 3 --no-cpu --extras none > gpurun_out/${tag}_c4wide_i8_c$c.json 2> gpurun_out/${tag}_c4wide_i8_c$c.err; tail -2 gpurun_out/${tag}_c4wide_i8_c$c.err
done
