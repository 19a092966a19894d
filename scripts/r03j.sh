for v in 0 1000 2000 3000 5000 8000; do
DLB_QF_OFFSET_NS=$v timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-cpu --extras none 2>/dev/null > gpurun_out/r03j_$v.json
python -c "
import json; d=json.load(open('gpurun_out/r03j_$v.json')); print('offset $v', d['value'], d['roofline']['phase_ms']['contract'], d['roofline']['frac'])"
done
