set -x
tag=$1
for v in 3 0; do
DLB_QF_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_qform.py -x -q 2>&1 | tail -2
for w in c4 c5solved; do
DLB_QF_VARIANT=$v timeout 300 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --extras none > gpurun_out/${tag}_bench_${w}_v$v.json 2> gpurun_out/${tag}_bench_${w}_v$v.err; tail -1 gpurun_out/${tag}_bench_${w}_v$v.err
done
done
