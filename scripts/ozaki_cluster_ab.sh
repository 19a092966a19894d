set -x
tag=$1
export DLB_OZAKI_VERBOSE=1
for c in 8 4 1; do
DLB_OZAKI_CLUSTER=$c timeout 300 python bench.py --workload c4wide_i8 --steps 10 --warmup 3 --no-cpu --extras none > gpurun_out/${tag}_c4wide_i8_c$c.json 2> gpurun_out/${tag}_c4wide_i8_c$c.err; grep -m3 "ozaki\|bench" gpurun_out/${tag}_c4wide_i8_c$c.err
done
