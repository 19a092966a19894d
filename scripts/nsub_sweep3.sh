run() { python bench.py --workload $1 --no-cpu --extras none --sustained 0 --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 nsub=${DLB_NSUB:-auto}', 'value %.3e e2e %.3e ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value']))"; }
python -m pytest tests/test_gpu_parity.py -x -q -k pipelined 2>&1 | tail -2
for w in c4 c5solved c5 c3; do run $w; done
