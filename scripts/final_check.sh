set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02final_pytest.log 2>&1; tail -2 gpurun_out/r02final_pytest.log
timeout 900 python bench.py > gpurun_out/r02final_bench_default.json 2> gpurun_out/r02final_bench_default.err; tail -2 gpurun_out/r02final_bench_default.err
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -1
