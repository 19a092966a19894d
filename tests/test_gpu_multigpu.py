"""Hardware check of the multi-GPU data plane: torchrun-launched, one process per GPU, NCCL for the plumbing,
results exchanged by peer stores from the finish kernel (csrc/gather.cu).  Needs >= 2 GPUs; on a 1-GPU box the single-process form of
the same kernels (world 1 window) is what runs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_gather_window_world1(gpu_required):
    import torch
    import desilike_b200
    from desilike_b200.distributed import PeerGather
    from oracle import configs
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True)
    plan = desilike_b200.Plan(spec, with_theory=False)
    N = 777
    X = configs.sample_points(spec, N, seed=3)
    ref = plan.eval(X)
    g = PeerGather(None, N * (1 + plan.ns))
    Xd = torch.as_tensor(X, device='cuda')
    for buf in (0, 1, 0):
        g.buffer(buf).zero_()
        ll = torch.empty(N, dtype=torch.float64, device='cuda')
        xs_local = torch.empty((N, plan.ns), dtype=torch.float64, device='cuda')
        plan.eval_gather(Xd, g, buf, 0, N, with_solved=True, out=dict(loglikelihood=ll, x_solved_plan_order=xs_local))
        g.wait(buf)
        torch.cuda.synchronize()
        g.check()
        w = g.buffer(buf).cpu().numpy()
        assert np.array_equal(w[:N], ref['logposterior'])
        assert np.array_equal(w[N:N + N * plan.ns].reshape(N, plan.ns)[:, plan._unperm], ref['x_solved'])
        assert np.array_equal(ll.cpu().numpy(), ref['loglikelihood'])
        assert np.array_equal(xs_local.cpu().numpy()[:, plan._unperm], ref['x_solved'])   # the rank's own solved values as a local output, too
    # a slice in the middle of a larger batch, no solved values
    g.buffer(0).zero_()
    plan.eval_gather(Xd[:100], g, 0, 300, 1000)
    g.wait(0)
    torch.cuda.synchronize()
    w = g.buffer(0).cpu().numpy()
    assert np.array_equal(w[300:400], ref['logposterior'][:100]) and not w[:300].any() and not w[400:1000].any()
    with pytest.raises(desilike_b200.DesilikeB200Error):   # batch larger than the window
        plan.eval_gather(Xd, g, 0, 0, 10**9)
    g.close()


def test_sharded_eval_matches_single_gpu(gpu_required):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip('needs >= 2 GPUs (run with gpurun --gpus 2)')
    world = 2 if ngpu < 4 else 4
    port = 29600 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'mgpu_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [line for line in res.stdout.splitlines() if line.startswith('MGPU_REPORT ')]
    assert lines, res.stdout[-2000:]
    report = json.loads(lines[-1][len('MGPU_REPORT '):])
    assert report['world'] == world and 'c5solved' in report['checks']
    assert report['transport'] in ('peer-store', 'nccl')
    print('multi-GPU transport:', report['transport'])
