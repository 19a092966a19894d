for v in 3 0 2 1; do
DLB_QF_VARIANT=$v timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu --extras none 2>/dev/null > gpurun_out/r03h_c4_v$v.json
python -c "
import json; d=json.load(open('gpurun_out/r03h_c4_v$v.json')); print('variant $v', d['value'], d['roofline']['phase_ms']['contract'], d['roofline']['frac'])"
done
