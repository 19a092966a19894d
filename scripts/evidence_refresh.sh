set -x
tag=r02y
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -1 gpurun_out/${tag}_bench_default.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:contract_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_full_contract_c4 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_contract_c4.err
