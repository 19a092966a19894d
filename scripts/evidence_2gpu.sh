set -x
tag=r02t
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --extras none > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err; tail -2 gpurun_out/${tag}_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_2gpu.json 2> gpurun_out/${tag}_bench_ref_2gpu.err; tail -2 gpurun_out/${tag}_bench_ref_2gpu.err; cat gpurun_out/${tag}_bench_ref_2gpu.json | cut -c1-300
