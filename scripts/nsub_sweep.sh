# host-buffer path: number of pipelined sub-batches (DLB_NSUB forces the count; unset = cost model of plan.cu choose_sub_batches)
run() { python bench.py --workload $1 --no-cpu --extras none --sustained 0 --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 nsub=${DLB_NSUB:-auto}', 'value %.3e e2e %.3e ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value']))"; }
for w in "$@"; do
  for ns in 1 2 4 6 8; do DLB_NSUB=$ns run $w; done
  run $w
done
