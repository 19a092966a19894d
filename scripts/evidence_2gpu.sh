set -x
tag=${1:-r02t}
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
DLB_GATHER_DMA=1 timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --extras none > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err; tail -2 gpurun_out/${tag}_bench_2gpu.err
DLB_GATHER_DMA=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 3 --extras none > gpurun_out/${tag}_bench_2gpu_nodma.json 2> gpurun_out/${tag}_bench_2gpu_nodma.err; tail -1 gpurun_out/${tag}_bench_2gpu_nodma.err
python - <<PY
import json
for f in ['gpurun_out/${tag}_bench_2gpu.json', 'gpurun_out/${tag}_bench_2gpu_nodma.json']:
    d = json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gather_check'))
PY
