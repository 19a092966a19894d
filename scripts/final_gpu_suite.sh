# full GPU suite + smoke (no bench)
tag=${1:-r02z}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -1
