set -x
tag=r02h
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2>&1
for spec in "c4 contract_kernel contract_c4" "c1 kaiser_kernel kaiser_c1" "c4wide_i8 ozaki_gemm_kernel ozaki_c4wide" "c3fft fftlog_kernel fftlog_c3"; do
  set -- $spec
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s 6 -c 1 -f -o gpurun_out/${tag}_ncu_full_$3 python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_$3.err
  ls -la gpurun_out/${tag}_ncu_full_$3.ncu-rep
done
