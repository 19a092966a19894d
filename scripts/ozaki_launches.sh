set -x
tag=$1
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload c4wide_i8 --steps 10 --warmup 3 --no-cpu --extras none > gpurun_out/${tag}_c4wide_i8.json 2> gpurun_out/${tag}_c4wide_i8.err; tail -3 gpurun_out/${tag}_c4wide_i8.err
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/${tag}_launches_c4wide_i8.csv python bench.py --workload c4wide_i8 --steps 1 --warmup 1 --no-cpu --extras none > /dev/null 2>&1
python - $tag <<'PY'
import csv, collections, sys
rows = list(csv.reader(open('gpurun_out/%s_launches_c4wide_i8.csv' % sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[i0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) <= iv: continue
    k = r[ik][:60]
    a = agg.setdefault(k, [0, 0.])
    a[0] += 1; a[1] += float(r[iv].replace(',', ''))
for k, (n, t) in agg.items(): print('%-62s n=%4d total=%10.1f us avg=%8.1f' % (k, n, t / 1e3, t / 1e3 / n))
PY
if [ "$2" = "ncu" ]; then
for spec in "ozaki_gemm_kernel<0> gemm0 0" "ozaki_gemm_kernel<1> gemm1 2"; do
  set -- $spec
  timeout 300 ncu --clock-control none --set full --import-source on -k "regex:ozaki_gemm_kernel" -s $3 -c 1 -f -o gpurun_out/${tag}_ncu_$2 python bench.py --workload c4wide_i8 --steps 1 --warmup 1 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_$2.err
done
fi
