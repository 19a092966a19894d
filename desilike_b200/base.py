"""Batched entry point: ``vmap`` with the reference's calling / return conventions (desilike/base.py:131-258, :232-383).

``vmap(likelihood, backend='cuda', errors='raise'|'return'|'nan', return_derived=False)`` returns ``wrapper(params)``
with ``params`` a dict name -> array (all of equal shape, rank >= 1).  The result is shaped ``params.shape`` (+ result
shape); with ``return_derived=True`` it is the tuple ``(logposterior, derived)`` where ``derived`` holds
``loglikelihood``, ``logprior`` and every analytically solved parameter (reference likelihoods/base.py:342-352,363-369).
With solved parameters, ``loglikelihood`` and ``logprior`` carry the reference's ``derivs`` tags (likelihoods/base.py:342-352,
:388-390, :409-411): their last axis is ``[value, H[i, j] for solved pairs i <= j]`` -- the likelihood / prior Hessian in the solved
parameters that ``Chain.sample_solved`` draws from -- addressed as ``array[()]`` and ``array[(name_i, name_j)]``.
``errors='return'`` -> ``(results, errs)`` with ``errs = {flat_index: (exception, traceback_str)}`` and failed slots filled
with a copy of the first good result; ``errors='nan'`` fills them with NaN (reference base.py:193-226).
"""
import numpy as np

from . import _capi
from .parameter import Samples, ParameterArray
from .likelihoods import PipelineError


class PointError(ArithmeticError):
    """Per-point failure reported by the CUDA path (status bitmask of include/desilike_b200.h)."""


def status_to_errors(status):
    """Per-point status bitmask -> {flat_index: (exception, traceback_str)}.  Non-finite theory is NOT an exception in the
    reference either (NaN propagates and samplers map it to -inf, samplers/base.py:187-191); a non-SPD normal matrix
    corresponds to the reference's LinAlgError in ``jnp.linalg.solve``."""
    errs = {}
    for idx in np.flatnonzero(np.asarray(status).ravel() & _capi.PT_NOT_SPD):
        errs[int(idx)] = (PointError('normal equations of the solved parameters are not positive definite'), '')
    return errs


def _check_params(params, known=None):
    """reference desilike/base.py:131-159."""
    params = dict(params)
    cshapes, names = [], list(params.keys())
    for name in names:
        array = params[name]
        if array is None: continue
        arr = np.asarray(array, dtype='f8')
        if not arr.shape:
            raise ValueError('input array {}: {} must be of rank >= 1 for lmap'.format(name, array))
        cshapes.append(arr.shape)
        params[name] = arr.ravel()
    if cshapes and not all(cshape == cshapes[0] for cshape in cshapes):
        raise ValueError('input shapes are different: {}'.format(dict(zip(names, cshapes))))
    return params, cshapes[0]


class ReferenceLikelihoodAdapter(object):
    """What ``vmap`` needs (``plan``, ``_points``) from one of the reference's own initialised likelihood objects: the plan is built from the
    spec extracted by ``adapters.spec_from_likelihood``, parameters absent from a call take the reference's current values."""

    def __init__(self, likelihood):
        from .adapters import spec_from_likelihood
        self.spec = spec_from_likelihood(likelihood)
        self._defaults = dict(self.spec['fiducial'])
        self._known = set(p.name for p in likelihood.all_params)
        self._plan = None

    @property
    def plan(self):
        if self._plan is None:
            from .plan import Plan
            self._plan = Plan(self.spec)
        return self._plan

    def _points(self, params):
        from .likelihoods import PipelineError
        for name in params:
            if name not in self._known: raise PipelineError('input parameter {} is not one of parameters: {}'.format(name, sorted(self._known)))
        names = self.spec['params']
        n = max([np.size(params[name]) for name in names if name in params] + [1])
        return np.column_stack([np.asarray(params[name], dtype='f8').ravel() if name in params else np.full(n, self._defaults[name]) for name in names])


def vmap(calculate, backend='cuda', errors='raise', return_derived=False, mpicomm=None, **kwargs):
    """Batched evaluation of a desilike_b200 likelihood.  ``backend`` must be 'cuda' (there is no CPU fallback);
    ``mpicomm`` may be a ``desilike_b200.distributed.Communicator`` to shard the batch over GPUs."""
    errors = str(errors)
    if backend != 'cuda':
        raise ValueError("desilike_b200.vmap only provides backend='cuda' (the reference's None / 'jax' / 'mpi' backends live in desilike itself)")
    if errors not in ('raise', 'return', 'nan'): raise ValueError("errors must be one of 'raise', 'return', 'nan'")
    likelihood = calculate
    if not hasattr(likelihood, '_points') and hasattr(likelihood, 'observables') and hasattr(likelihood, 'all_params'):
        likelihood = ReferenceLikelihoodAdapter(likelihood)   # one of desilike's OWN likelihood objects: extract its constants into a plan

    def wrapper(params, **kw):
        rd = kw.pop('return_derived', return_derived)
        params, shape = _check_params(params)
        X = likelihood._points(params)
        plan = likelihood.plan
        with_hessian = bool(rd and plan.ns and mpicomm is None)
        if mpicomm is not None:
            res = mpicomm.eval_sharded(plan, X)   # (the per-point Hessians are not exchanged: sharded calls return untagged derived values)
        else:
            res = plan.eval(X, return_hessian=with_hessian)
        errs = status_to_errors(res['status'])
        if errs and errors == 'raise':
            raise errs[min(errs)][0]
        out = {key: np.array(res[key]) for key in ['logposterior', 'loglikelihood', 'logprior']}
        xs = np.array(res['x_solved']) if 'x_solved' in res else None
        good = []
        if errs:
            good = [i for i in range(X.shape[0]) if i not in errs]
            for key in out:
                fill = np.nan if (errors == 'nan' or not good) else out[key][good[0]]
                out[key][list(errs)] = fill
            if xs is not None:
                xs[list(errs)] = np.nan if (errors == 'nan' or not good) else xs[good[0]]
        results = out['logposterior'].reshape(shape)
        if rd and with_hessian:
            # derivs tags: () then every pair of solved parameters exported to derived (reference likelihoods/base.py:342-352)
            solved = [s for s in plan.spec.get('solved', [])]
            keep = [i for i, s in enumerate(solved) if not str(s.get('derived', '')).endswith('not_derived')]
            pairs = [(i, j) for a, i in enumerate(keep) for j in keep[a:]]
            tags = [()] + [(solved[i]['name'], solved[j]['name']) for i, j in pairs]
            prec = np.array([1. / float(s['prior']['scale'])**2 if (s.get('prior') or {}).get('dist') == 'norm' and np.isfinite(float(s['prior'].get('scale', np.inf))) else 0.
                             for s in solved])
            H = np.array(res['hessian'])                       # posterior Hessian in the solved parameters, spec order
            if errs: H[list(errs)] = np.nan if (errors == 'nan' or not good) else H[good[0]]
            ii, jj = [i for i, _ in pairs], [j for _, j in pairs]
            Hprior = -np.diag(prec)
            ll = np.concatenate([out['loglikelihood'][:, None], (H - Hprior)[:, ii, jj]], axis=-1)
            lp = np.concatenate([out['logprior'][:, None], np.broadcast_to(Hprior[ii, jj], (X.shape[0], len(pairs)))], axis=-1)
            derived = Samples(loglikelihood=ParameterArray(ll.reshape(shape + (len(tags),)), derivs=tags),
                              logprior=ParameterArray(lp.reshape(shape + (len(tags),)), derivs=tags))
            for i, name in enumerate(plan.solved_names):
                derived[name] = xs[:, i].reshape(shape)
            results = (results, derived)
        elif rd:
            derived = Samples(loglikelihood=out['loglikelihood'].reshape(shape), logprior=out['logprior'].reshape(shape))
            for i, name in enumerate(likelihood.plan.solved_names):
                derived[name] = xs[:, i].reshape(shape)
            results = (results, derived)
        if errors == 'return':
            return results, errs
        return results

    wrapper.__wrapped__vmap__ = calculate
    wrapper.__wrapped__errors__ = errors
    return wrapper
