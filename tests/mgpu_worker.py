"""Worker of tests/test_gpu_multigpu.py (launched by torch.distributed.run, one process per GPU): the point-sharded evaluation through
the public multi-GPU API must reproduce a single-GPU evaluation of the full batch on every rank -- on the hardware data plane
(peer stores over NVLink, or the NCCL fallback), not on a stand-in."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import desilike_b200
    from desilike_b200.distributed import Communicator, PeerGather, local_slice
    from oracle import configs
    rank, world, local_rank = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    comm = Communicator()
    report = dict(rank=rank, world=world, checks=[])
    for name, spec, N in [('c5solved', configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True), 1001),
                          ('c2', configs.config_c2(nkin=50, nbins=12, order=2), 5),
                          ('c2_fewer_points_than_ranks', configs.config_c2(nkin=50, nbins=12, order=2), 1)]:
        plan = desilike_b200.Plan(spec, with_theory=False)
        X = configs.sample_points(spec, N, seed=11)
        X[0, 0] = 5.   # outside the prior: -inf must travel too
        ref = plan.eval(X)
        for trial in range(3):   # alternates the window buffers
            out = comm.eval_sharded(plan, X)
            for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
                if key not in ref: continue
                assert np.array_equal(np.asarray(out[key]), np.asarray(ref[key]), equal_nan=True) if key != 'status' else np.array_equal(out[key], ref[key]), (name, key, trial)
        Xd = torch.as_tensor(X, device=dev)
        outd = comm.eval_sharded(plan, Xd, return_all=False)
        torch.cuda.synchronize()
        assert np.array_equal(outd['logposterior'].cpu().numpy(), ref['logposterior'])
        if plan.ns: assert np.array_equal(outd['x_solved'].cpu().numpy(), ref['x_solved'])
        report['checks'].append(name)
        plan.close()
    report['transport'] = comm.transport
    # generic publish / wait on a dedicated window
    if comm.transport == 'peer-store':
        g = PeerGather(comm, 64 * world, nbuf=1)
        for it in range(4):
            g.publish(0, torch.full((64,), float(100 * it + rank), dtype=torch.float64, device=dev), 64 * rank)
            g.wait(0)
            got = g.buffer(0).clone()
            torch.cuda.synchronize()
            g.check()
            assert np.array_equal(got.cpu().numpy(), np.repeat(100. * it + np.arange(world), 64)), it
            dist.barrier()   # nbuf = 1: nobody may publish iteration it + 1 before everybody has read iteration it
        g.close()
        report['checks'].append('publish')
    comm.close()
    dist.barrier()
    if rank == 0:
        print('MGPU_REPORT ' + json.dumps(report))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
