# C2 (65 536 points) through the host-buffer path with different sub-chunk floors of the H2D | kernel | D2H pipeline
for sm in 65536 32768 16384 8192; do
  DLB_SUB_MIN=$sm python bench.py --workload c2 --no-cpu --extras none --sustained 0 --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('submin', $sm, 'value %.3e e2e %.3e ratio %.2f ms %.4f' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value'], d['ms_per_step']))"
done
