run() { python bench.py --workload $1 --no-cpu --extras none --sustained 0 --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 nsub=${DLB_NSUB:-auto}', 'value %.3e e2e %.3e ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value']))"; }
for ns in 3 5 7; do DLB_NSUB=$ns run c4; done
run c4; run c5
DLB_NSUB=1 run c5solved; run c5solved
DLB_NSUB=1 run c3; run c3
run c2
