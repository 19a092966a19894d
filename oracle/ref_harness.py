"""Run the reference's OWN NumPy code path (read-only /root/reference) under stub third-party modules.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, smoke() or bench.py imports this module.  It is used by tests/golden/make_golden.py to generate the
committed fixtures that pin the oracle, and by tests/test_oracle_vs_reference.py
(skipped when /root/reference is absent).

The reference classes are used UNMODIFIED; only the absent third-party packages (cosmoprimo, lsstypes) are replaced
by the shells in oracle/stubs/ whose fiducial cosmology is the closed-form oracle/synthetic.py.
"""
import os
import sys
import warnings

import numpy as np

REFERENCE = '/root/reference'
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'desilike'))


def setup():
    """Put stubs, then the reference, on sys.path (idempotent)."""
    if not available():
        raise RuntimeError('reference not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in [root, REFERENCE, _STUBS]:
        if path in sys.path: sys.path.remove(path)
        sys.path.insert(0, path)
    warnings.filterwarnings('ignore')
    import desilike  # noqa: F401
    return desilike


def _set_params(calculator, spec, extra_fixed=()):
    """Make the reference pipeline vary exactly spec['params'] with the spec priors; fix everything else."""
    for param in calculator.all_params:
        name = param.name
        if name in spec['params']:
            prior = spec['priors'].get(name)
            kw = dict(fixed=False)
            if prior is not None:
                kw['prior'] = {k: v for k, v in prior.items()}
            param.update(**kw)
        elif not param.derived or param.solved:
            pass
    return calculator


def reference_kaiser_likelihood(spec, eftlike=False, solved_mode=None):
    """Reference likelihood for a single-observable Kaiser(+ShapeFit) spec (C1-like, optional window)."""
    setup()
    from desilike.theories.galaxy_clustering import (ShapeFitPowerSpectrumTemplate, KaiserTracerPowerSpectrumMultipoles,
                                                     EFTLikeKaiserTracerPowerSpectrumMultipoles)
    from desilike.observables.galaxy_clustering import TracerPowerSpectrumMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    obs = spec['observables'][0]
    th = obs['theory']
    template = ShapeFitPowerSpectrumTemplate(z=1., a=th['template']['a'], kp=th['template']['kp'], apmode=th['template']['apmode'])
    cls = EFTLikeKaiserTracerPowerSpectrumMultipoles if eftlike else KaiserTracerPowerSpectrumMultipoles
    theory = cls(template=template, mu=th["nmu"])
    nk = len(th['k'])
    nell = len(th['ells'])
    kw = {}
    if obs.get('wmatrix') is not None:
        kout = obs['kout']
        kw = dict(wmatrix=np.asarray(obs['wmatrix_raw']), kin=th['k'], ellsin=th['ells'], k=[kout] * nell, ells=th['ells'], shotnoise=obs.get('shotnoise'))
    else:
        kw = dict(k=[th['k']] * nell, ells=th['ells'])
    if obs.get('systematic_templates') is not None: kw['systematic_templates'] = [np.asarray(t) for t in obs['systematic_templates']]
    observable = TracerPowerSpectrumMultipolesObservable(data=np.asarray(obs['data']), theory=theory, transform=obs.get('transform'), **kw)
    precision = spec['precision'][0] if isinstance(spec['precision'], (list, tuple)) else spec['precision']
    likelihood = ObservablesGaussianLikelihood(observables=[observable], precision=np.array(precision))
    likelihood.init.params = likelihood.init.params  # touch
    for param in likelihood.all_params:
        name = param.name
        if name in spec['params']:
            prior = spec['priors'].get(name)
            param.update(fixed=False, prior=dict(prior) if prior is not None else None)
        elif any(s['name'] == name for s in spec.get('solved', [])):
            s = [s for s in spec['solved'] if s['name'] == name][0]
            param.update(fixed=False, value=s['value'], prior=dict(s['prior']) if s.get('prior') else None,
                         derived=solved_mode or ('.marg' if s.get('marg', True) else '.best'))
        elif not param.derived:
            param.update(fixed=True)
    return likelihood


def run_vmap(likelihood, spec, X, return_theory=True):
    """Serial ``vmap`` of the reference (reference desilike/base.py:232-273) -> dict of arrays."""
    from desilike.base import vmap
    params = {name: X[:, i] for i, name in enumerate(spec['params'])}
    out = dict(logposterior=np.empty(X.shape[0]), loglikelihood=np.empty(X.shape[0]), logprior=np.empty(X.shape[0]))
    theory = []

    def calculate(p):
        val = likelihood(p)
        theory.append(np.array(likelihood.flattheory))
        return np.array([val, float(np.ravel(likelihood.loglikelihood)[0]), float(np.ravel(likelihood.logprior)[0])])

    res = vmap(calculate, backend=None, errors='raise')(params)
    out['logposterior'], out['loglikelihood'], out['logprior'] = res[:, 0], res[:, 1], res[:, 2]
    if return_theory:
        out['flattheory'] = np.array(theory)
    return out


def install_jacfwd_stub(step=1.):
    """The reference's FastFisher.__call__ (reference desilike/likelihoods/base.py:129-200) does ``import jax`` and calls
    ``jax.jacfwd(getter)(x)``; jax is absent.  The theory is exactly linear in the solved parameters, so a central
    difference with unit step IS the Jacobian (up to rounding).  Injecting this shim AFTER desilike was imported
    (so desilike.jax keeps its NumPy fallback) lets the reference's own FastFisher/_solve code run unmodified.
    ``step`` < 1: for a flatdiff that is NOT linear in the solved parameters (observable transform='cubic') the shim returns the Richardson-
    extrapolated central difference (4 D(step / 2) - D(step)) / 3, accurate to O(step^4) -- far below the fixture tolerances."""
    import types

    def jacfwd(fun, argnums=0, has_aux=False, holomorphic=False):
        def jac(x):
            x = np.asarray(x, dtype='f8')
            cols = []
            for i in range(x.size):
                e = np.zeros_like(x); e[i] = step

                def central(h):
                    up = [np.array(a) for a in fun(x + h * e)]
                    dn = [np.array(a) for a in fun(x - h * e)]
                    return [(u - d) / (2. * h * step) for u, d in zip(up, dn)]

                if step == 1.: cols.append(central(1.))
                else: cols.append([(4. * a - b) / 3. for a, b in zip(central(0.5), central(1.))])
            return [np.stack([c[j] for c in cols], axis=-1) for j in range(len(cols[0]))]
        return jac

    mod = types.ModuleType('jax')
    mod.jacfwd = jacfwd
    sys.modules['jax'] = mod
