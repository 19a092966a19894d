# 2 GPUs: data-plane test + one short bench line (host-buffer e2e leg goes through the sub-batch pipeline with peer stores per sub-batch)
tag=${1:-r02z}
timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --extras none --no-cpu --sustained 0 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err; tail -2 gpurun_out/${tag}_bench_2gpu.err
python - <<PY
import json
d = json.load(open('gpurun_out/${tag}_bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gather_check'))
PY
