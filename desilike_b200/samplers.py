"""Device-resident ensemble sampler around a :class:`~desilike_b200.plan.Plan`.

What it replaces: the reference's ``EmceeSampler`` (desilike/samplers/emcee.py:17-110) builds
``emcee.EnsembleSampler(nwalkers, ndim, self.logposterior, vectorize=True)`` (:68) and iterates it in ``_run_one`` (:100-110), every
iteration calling the vectorised ``logposterior(values[N, nparams])`` of desilike/samplers/base.py:144-200 on the host.  Here the
walkers, their log-posteriors and the proposals stay in HBM: each iteration is ``dlb_ensemble_step`` (propose kernel -> batched
likelihood -> accept kernel, csrc/sampler.cu), enqueued on the current CUDA stream with no host round trip.

Multi-GPU: one process per GPU, each evolving its own sub-ensemble (partners are drawn from the local complementary half, so no
walker positions cross NVLink); the only collective is the all-gather of the log-posteriors
(:meth:`EnsembleSampler.gather_logposterior`), as the reference's ``mpicomm.bcast(logposterior)`` (samplers/base.py:200).
"""
import ctypes

import numpy as np

from . import _capi


class EnsembleSampler(object):
    """
    Affine-invariant ensemble sampler (stretch move, red-blue split) with walkers resident on the GPU.

    Parameters
    ----------
    plan : Plan
        Evaluation plan of the likelihood (parameters ordered as ``plan.params``).
    nwalkers : int
        Number of walkers ON THIS RANK (even, >= 2; the reference default is ``2 * max((int(2.5 * ndim) + 1) // 2, 2)``,
        samplers/emcee.py:65).
    seed : int
        Key of the counter-based random stream.
    a : float
        Stretch scale (emcee default 2).
    rank_offset : int
        Global index of this rank's first walker within a half (distinct per rank when an ensemble is sharded).
    """

    def __init__(self, plan, nwalkers, seed=42, a=2., rank_offset=0):
        import torch
        self._torch = torch
        self.plan = plan
        self.nwalkers, self.ndim = int(nwalkers), plan.d
        if self.nwalkers < 2 or self.nwalkers % 2:
            raise ValueError('nwalkers must be even and >= 2')
        self.seed, self.a, self.rank_offset = int(seed), float(a), int(rank_offset)
        self.iteration = 0
        self.walkers = self.logposterior = None
        self._naccept = None
        self._lib = _capi.load()

    def set_start(self, start):
        """Initial walker positions [nwalkers, ndim] (NumPy or CUDA tensor); evaluates their log-posteriors."""
        torch = self._torch
        X = torch.as_tensor(np.ascontiguousarray(start, dtype=np.float64) if not hasattr(start, 'is_cuda') else start, dtype=torch.float64, device='cuda').contiguous()
        if X.shape != (self.nwalkers, self.ndim):
            raise ValueError('expected start of shape ({:d}, {:d})'.format(self.nwalkers, self.ndim))
        self.walkers = X.clone()
        res = self.plan.eval(self.walkers, return_solved=False, return_status=False)
        self.logposterior = res['logposterior']
        if not bool(torch.isfinite(self.logposterior).all()):
            raise ValueError('initial walkers must have finite log-posterior')
        self._naccept = torch.zeros(1, dtype=torch.int64, device=X.device)
        self.iteration = 0

    def step(self):
        """One iteration (both halves), enqueued on the current stream."""
        torch = self._torch
        stream = torch.cuda.current_stream(self.walkers.device).cuda_stream
        _capi.check(self._lib.dlb_ensemble_step(self.plan._handle, self.walkers.data_ptr(), self.logposterior.data_ptr(), self.nwalkers,
                                                self.seed, self.iteration, self.a, self.rank_offset, self._naccept.data_ptr(), ctypes.c_void_p(stream)))
        self.iteration += 1

    def run(self, niterations, thin_by=1, store=True):
        """Run ``niterations``; returns (chain [nstored, nwalkers, ndim], logposterior [nstored, nwalkers]) as CUDA tensors when ``store``."""
        torch = self._torch
        if self.walkers is None:
            raise ValueError('call set_start first')
        chain, logp = [], []
        for it in range(niterations):
            self.step()
            if store and (it + 1) % thin_by == 0:
                chain.append(self.walkers.clone()); logp.append(self.logposterior.clone())
        if store and chain:
            return torch.stack(chain), torch.stack(logp)
        return None, None

    @property
    def acceptance_fraction(self):
        if not self.iteration: return 0.
        return float(self._naccept.item()) / (self.iteration * self.nwalkers)

    def gather_logposterior(self):
        """All-gather of the current log-posteriors over torch.distributed ranks (NCCL over NVLink): [world, nwalkers]."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self.logposterior[None, :]
        out = self._torch.empty((dist.get_world_size(), self.nwalkers), dtype=self._torch.float64, device=self.logposterior.device)
        dist.all_gather_into_tensor(out, self.logposterior)
        return out
