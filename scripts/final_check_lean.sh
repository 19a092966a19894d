# full GPU suite + the headline bench line (no extras) + smoke
tag=${1:-r02z}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --extras none > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; tail -1 gpurun_out/${tag}_bench_c4.err
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -1
