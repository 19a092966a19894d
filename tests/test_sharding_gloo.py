"""Multi-process host logic of the point-sharded path on CPU: world_size-2 gloo, no GPU (the per-rank evaluation is replaced by
the oracle -- only the partition rule and the ragged all-gather are under test)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from desilike_b200.distributed import Communicator, local_slice
    from oracle import configs, pipeline
    spec = configs.config_c5(ntracers=2, nkin=20, nbins=6, order=1, solved=True)
    X = configs.sample_points(spec, N, seed=7)
    comm = Communicator()

    def evaluate(Xl):
        res = pipeline.evaluate(spec, Xl)
        res['status'] = np.full(len(Xl), rank, dtype=np.int32)
        return res

    out = comm.eval_sharded(None, X, evaluate=evaluate)
    np.savez(os.path.join(tmpdir, 'rank{:d}.npz'.format(rank)), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize('N', [11, 2])
def test_world_size_2_gloo(tmp_path, N):
    import torch.multiprocessing as mp
    from desilike_b200.distributed import local_size, local_slice
    assert [local_size(11, r, 4) for r in range(4)] == [3, 3, 3, 2]
    assert local_slice(11, 3, 4) == slice(9, 11)
    port = 29500 + (os.getpid() + N) % 2000
    mp.spawn(_worker, args=(2, port, N, str(tmp_path)), nprocs=2, join=True)
    from oracle import configs, pipeline
    spec = configs.config_c5(ntracers=2, nkin=20, nbins=6, order=1, solved=True)
    X = configs.sample_points(spec, N, seed=7)
    ref = pipeline.evaluate(spec, X)
    outs = [np.load(os.path.join(str(tmp_path), 'rank{:d}.npz'.format(r))) for r in range(2)]
    for out in outs:
        assert np.allclose(out['logposterior'], ref['logposterior'], rtol=1e-12, atol=1e-10)   # (BLAS blocking differs with the slice size)
        assert np.allclose(out['x_solved'], ref['x_solved'], rtol=1e-10, atol=1e-10)
        assert list(out['status']) == [0] * local_size(N, 0, 2) + [1] * local_size(N, 1, 2)
    assert np.array_equal(outs[0]['logposterior'], outs[1]['logposterior'])
