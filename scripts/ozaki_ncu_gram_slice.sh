set -x
tag=$1
for spec in "ozaki_gram_kernel gram 0" "ozaki_slice_kernel slice 1" "ozaki_slice_kernel slice3 3"; do
  set -- $spec
  timeout 300 ncu --clock-control none --set full --import-source on -k "regex:$1" -s $3 -c 1 -f -o gpurun_out/${tag}_ncu_$2 python bench.py --workload c4wide_i8 --steps 1 --warmup 1 --no-cpu --extras none > /dev/null 2> gpurun_out/${tag}_ncu_$2.err
done
