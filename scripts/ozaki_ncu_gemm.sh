set -x
tag=$1
export DLB_OZAKI_CLUSTER=${2:-1}
timeout 300 ncu --clock-control none --set full --import-source on -k "regex:ozaki_gemm_kernel" -s 15 -c 1 -f -o gpurun_out/${tag}_ncu_gemmc python bench.py --workload c4wide_i8 --steps 1 --warmup 1 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_gemmc.err
