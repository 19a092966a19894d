set -x
timeout 600 python -m pytest tests/test_gpu_qform.py tests/test_gpu_parity.py tests/test_gpu_api.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu --extras none 2>&1 >gpurun_out/r03g_c4.json | tail -1
timeout 300 python bench.py --workload c5solved --steps 50 --warmup 5 --no-cpu --extras none 2>&1 >gpurun_out/r03g_c5s.json | tail -1
python - <<'PY'
import json
for f in ['gpurun_out/r03g_c4.json','gpurun_out/r03g_c5s.json']:
    d=json.load(open(f)); print(f, d['value'], d['roofline']['phase_ms'], d['roofline']['frac'])
PY
