"""PyTorch custom-op layer over the C-ABI:

    pid = desilike_b200.torch_ops.register_plan(plan)
    logpost, loglike, logprior, x_solved, status = torch.ops.desilike_b200.eval_batch(pid, params)   # params: f64 [N, d] CUDA tensor

The op runs ``dlb_eval_batch`` on the current CUDA stream of ``params``' device, in place on device memory (no host round trip); it has a
fake (meta) implementation, so it composes with ``torch.compile`` / FakeTensor tracing of a sampler step.  PyTorch is plumbing here (device
memory, streams): the arithmetic is the hand-written sm_100a kernels behind the C-ABI.  ``x_solved`` is in the plan's solved order
('.marg' parameters first); ``plan.solved_names`` / ``Plan.eval`` give the spec order.
"""
import itertools
from typing import Tuple

import torch

_PLANS = {}
_IDS = itertools.count(1)


def register_plan(plan, d=None, ns=None):
    """-> integer handle for ``torch.ops.desilike_b200.eval_batch``.  ``plan`` may be None with explicit (d, ns) for shape tracing only."""
    pid = next(_IDS)
    _PLANS[pid] = (plan, int(plan.d if plan is not None else d), int(plan.ns if plan is not None else ns))
    return pid


def unregister_plan(pid):
    _PLANS.pop(int(pid), None)


def _lookup(pid):
    try:
        return _PLANS[int(pid)]
    except KeyError:
        raise ValueError('unknown plan handle {} (use desilike_b200.torch_ops.register_plan)'.format(pid))


@torch.library.custom_op('desilike_b200::eval_batch', mutates_args=(), device_types='cuda')
def eval_batch(plan_id: int, params: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    plan, d, ns = _lookup(plan_id)
    if plan is None: raise RuntimeError('plan handle {} was registered for shape tracing only'.format(plan_id))
    out = plan._eval_torch(params.contiguous(), True, True, None)
    N = params.shape[0]
    xs = out['x_solved_plan_order'] if ns else torch.empty((N, 0), dtype=torch.float64, device=params.device)
    return out['logposterior'], out['loglikelihood'], out['logprior'], xs, out['status']


@eval_batch.register_fake
def _(plan_id, params):
    _, d, ns = _lookup(plan_id)
    if params.dim() != 2 or params.shape[1] != d: raise ValueError('expected params of shape (N, {:d}), got {}'.format(d, tuple(params.shape)))
    N = params.shape[0]
    f = lambda *shape: params.new_empty(shape, dtype=torch.float64)
    return f(N), f(N), f(N), f(N, ns), params.new_empty((N,), dtype=torch.int32)
