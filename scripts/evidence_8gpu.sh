set -x
tag=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/${tag}_bench_8gpu.json 2> gpurun_out/${tag}_bench_8gpu.err; tail -3 gpurun_out/${tag}_bench_8gpu.err
if [ "$2" = "ab" ]; then
DLB_GATHER_DMA=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 50 --warmup 3 --extras none > gpurun_out/${tag}_bench_8gpu_nodma.json 2> gpurun_out/${tag}_bench_8gpu_nodma.err; tail -1 gpurun_out/${tag}_bench_8gpu_nodma.err
fi
