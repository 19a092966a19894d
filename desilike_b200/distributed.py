"""Multi-GPU sharding of the batch by parameter point (one process per GPU, ``torch.distributed``).

The reference's only parallelism on this path is data-parallel over points: ``vmap(backend='mpi')`` scatters chunks with
``mpi.scatter`` (desilike/base.py:310-316, partition rule ``mpi.local_size`` desilike/mpi.py:145-150), each rank evaluates its
slice, and results come back by a pickled ``reduce`` + ``bcast`` (base.py:335; samplers/base.py:196-200).  Here: contiguous
slices by the same rule, each rank runs its slice through its own device-resident plan, and ONE all-gather returns
the log-posteriors (and, on request, the solved parameters) to every rank -- NCCL over NVLink on GPUs, gloo on CPU for
the host-logic tests.  No exchange happens inside the evaluation.
"""
import numpy as np


def local_size(size, rank, world):
    """Number of points of rank ``rank`` (reference desilike/mpi.py:145-150: the first ``size % world`` ranks take one more)."""
    return size // world + (1 if rank < size % world else 0)


def local_slice(size, rank, world):
    start = sum(local_size(size, r, world) for r in range(rank))
    return slice(start, start + local_size(size, rank, world))


class Communicator(object):
    """Thin wrapper over an initialised ``torch.distributed`` process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised (launch with torchrun, then dist.init_process_group)')
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def allgather_ragged(self, local, total, device=None):
        """All-gather 1-D / 2-D float64 chunks of unequal length (padded to the largest chunk)."""
        import torch
        local = torch.as_tensor(local)
        trailing = tuple(local.shape[1:])
        nmax = local_size(total, 0, self.size)
        buf = torch.zeros((nmax,) + trailing, dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
        out = torch.empty((self.size * nmax,) + trailing, dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.reshape((self.size, nmax) + trailing)
        return torch.cat([out[r, :local_size(total, r, self.size)] for r in range(self.size)], dim=0)

    def eval_sharded(self, plan, X, evaluate=None):
        """Every rank passes the FULL batch X [N, d]; returns the full result dict on every rank."""
        import torch
        X = np.ascontiguousarray(X, dtype=np.float64)
        N = X.shape[0]
        sl = local_slice(N, self.rank, self.size)
        res = (evaluate or plan.eval)(X[sl])
        device = torch.device('cuda', torch.cuda.current_device()) if self.backend == 'nccl' else torch.device('cpu')
        out = {}
        for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
            if key not in res: continue
            local = torch.as_tensor(np.asarray(res[key]), device=device)
            if local.ndim == 2 and local.shape[1] == 0:
                out[key] = np.zeros((N, 0)); continue
            out[key] = self.allgather_ragged(local, N).cpu().numpy()
        return out
