#!/bin/bash
# usage (on the GPU box, through gpurun): scripts/profile_suite.sh <tag> bench|ncu
# bench: bench lines of every workload + the default line (with the CPU baseline) + the reference arm -> gpurun_out/<tag>_*.json
# ncu:   launch list of the default bench + `ncu --set full` captures of the dominant kernels -> gpurun_out/<tag>_*.{csv,ncu-rep}
tag=$1; what=$2
if [ "$what" = "bench" ]; then
  timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
  scripts/bench_all.sh $tag c1 c2 c3 c3fft c4 c5 c5solved
  timeout 300 python bench.py --workload c5step --steps 20 --warmup 3 > gpurun_out/${tag}_bench_c5step.json 2> gpurun_out/${tag}_bench_c5step.err
  python -c "
import json
for n in ['default', 'reference_arm', 'c5step']:
    d = json.load(open('gpurun_out/${tag}_bench_%s.json' % n)); print(n, d['value'], d['e2e']['value'], d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))"
else
  NCU="ncu --clock-control none"
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
  for spec in "c2 contract_kernel contract_c2" "c5 contract_kernel contract_c5" "c4 contract_kernel contract_c4" "c4 finish_solve_thread_kernel solve_c4" "c3 bao_kernel bao_c3" "c1 kaiser_kernel kaiser_c1"; do
    set -- $spec
    timeout 600 $NCU --set full --import-source on -k regex:$2 -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_full_$3 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/${tag}_ncu_$3.err
    ls -la gpurun_out/${tag}_ncu_full_$3.ncu-rep
  done
fi
