set -x
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; tail -3 gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:contract_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_full_contract_c4 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_contract_c4.err
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')" 2>&1 | tail -2
