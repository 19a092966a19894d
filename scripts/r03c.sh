set -x
timeout 600 python -m pytest tests/test_gpu_api.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 30 --warmup 3 --extras c2,c5solved,c1,c3 > gpurun_out/r03c_bench.json 2> gpurun_out/r03c_bench.err; tail -6 gpurun_out/r03c_bench.err
