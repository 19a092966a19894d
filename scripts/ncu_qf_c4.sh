set -x
tag=$1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:contract_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_ncu_qf_c4 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_qf_c4.err
