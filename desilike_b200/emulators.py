"""Emulator files on the batched path: reader / writer of the reference's emulator ``.npy`` state and the
``EmulatedCalculator`` that plugs an emulated perturbation-theory calculator into a tracer theory.

Reference API kept (desilike/emulators/__init__.py): ``Emulator.save(fn)`` = ``np.save(fn, state, allow_pickle=True)`` (:240-248),
``EmulatedCalculator.load(fn)`` (:411-415), ``emulator.to_calculator()`` (:150-208).  State keys written by desilike itself
(:210-216): varied_params, in_calculator_state, calculator__class__, is_calculator_sequence, yaml_data, all_params; keys
of the (absent, third-party) cosmoprimo base class as evidenced by desilike/emulators/conversion.py:48-54,102-104 and the
back-compat shim :218-235: engines {name: {name: 'taylor'|'mlp', params, xshape, yshape, xoperations, yoperations,
center/derivatives/powers | model_operations/model_yoperations}}, xoperations, yoperations, defaults, fixed.
Operations are {expression, inverse, locals} strings (conversion.py:24-33,76-84); they are NOT eval'ed here: a whitelist of
the expressions the reference itself emits is pattern-matched and lowered to the CUDA MLP generator; anything else raises.
The schema beyond those call sites is third-party, so compatibility is best effort and round-trip tested with this writer.

Fitting: ``Emulator.fit_taylor`` builds a Taylor emulator from GPU theory evaluations on a finite-difference stencil;
MLP training is not part of this module.
"""
import os
import re

import numpy as np

from .parameter import Parameter, ParameterCollection
from .theories import BaseCalculator

_DENSE = '(v[..., None, :] @ kernel)[..., 0, :] + bias'
_ACTS = {'v / (1 + jnp.exp(-v))': 'silu', 'jnp.maximum(v, 0.)': 'relu', 'jnp.tanh(v)': 'tanh'}
_ACT_EXPR = {v: k for k, v in _ACTS.items()}
_MINMAX = '(v - limits[..., 0]) / (limits[..., 1] - limits[..., 0])'
_MINMAX_INV = 'v * (limits[..., 1] - limits[..., 0]) + limits[..., 0]'
_PCA = 'v @ eigvecs + mean'


def _norm(expr):
    expr = re.sub(r'\s+', ' ', str(expr).strip())
    while expr.startswith('(') and expr.endswith(')') and _balanced(expr[1:-1]): expr = expr[1:-1].strip()
    return expr


def _balanced(s):
    depth = 0
    for c in s:
        depth += (c == '(') - (c == ')')
        if depth < 0: return False
    return depth == 0


def _op(expression, inverse=None, locals=None):
    return {'expression': expression, 'inverse': inverse, 'locals': dict(locals or {})}


def _op_fields(op):
    get = lambda *keys: next((op[k] for k in keys if k in op), None)
    return get('expression', '_expression', 'direct'), get('inverse', '_inverse'), get('locals', '_locals') or {}


class Emulator(object):
    """Container of an emulator state (no fitting): ``Emulator.load(fn)``, ``.save(fn)``, ``.to_calculator()``."""

    def __init__(self, state):
        self.state = state
        for key in ['engines', 'fixed']:
            if key not in state: raise ValueError('emulator state has no {!r} entry'.format(key))

    @classmethod
    def load(cls, filename):
        state = np.load(filename, allow_pickle=True)[()]
        return cls(state)

    def save(self, filename, yaml=False):
        dirname = os.path.dirname(str(filename))
        if dirname: os.makedirs(dirname, exist_ok=True)
        np.save(filename, self.state, allow_pickle=True)

    def to_calculator(self):
        return EmulatedCalculator(self)

    # ---- writers ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _base_state(params, defaults, fixed, names, calculator_class, all_params):
        all_params = all_params or [Parameter(p, value=defaults.get(p, 0.)) for p in params]
        return {'engines': {}, 'xoperations': [], 'yoperations': [], 'defaults': dict(defaults), 'fixed': dict(fixed, names=list(names)),
                'varied_params': list(params), 'in_calculator_state': list(names), 'calculator__class__': tuple(calculator_class),
                'is_calculator_sequence': False, 'yaml_data': {'class': calculator_class[1], 'info': {}, 'params': {}},
                'all_params': {'data': [dict(basename=p.basename, namespace=p.namespace, value=p.value, fixed=p.fixed, prior=p.prior, ref=p.ref, derived=p.derived)
                                        for p in all_params]}}

    @classmethod
    def from_taylor(cls, params, center, powers, tables, fixed, defaults=None, all_params=None,
                    calculator_class=('desilike.theories.galaxy_clustering.full_shape', 'KaiserPowerSpectrumMultipoles')):
        """``tables``: {name: derivatives [T, *yshape]} in the layout of the reference's fit (emulators/__init__.py:471-507)."""
        defaults = dict(zip(params, np.asarray(center, dtype='f8'))) if defaults is None else defaults
        state = cls._base_state(params, defaults, fixed, list(tables), calculator_class, all_params)
        for name, deriv in tables.items():
            deriv = np.asarray(deriv, dtype='f8')
            state['engines'][name] = {'name': 'taylor', 'params': list(params), 'xshape': (len(params),), 'yshape': tuple(deriv.shape[1:]),
                                      'xoperations': [], 'yoperations': [], 'center': np.asarray(center, dtype='f8'),
                                      'derivatives': deriv, 'powers': np.asarray(powers, dtype='i4')}
        return cls(state)

    @staticmethod
    def _theory_tables(theory, names):
        """-> (evaluate(X) -> {table: [n, ny]}, {table: yshape}, fixed, calculator_class): the emulated tables of a direct theory, evaluated ON THE GPU
        through its feature generator (``dlb_features_batch``).  Kaiser tracer theories: the perturbation-theory tables pk_dd / pk_dt / pk_tt
        (what the reference emulates as ``theory.pt``); damped-BAO tracer theories: the multipoles without broadband, 'power' [nell, nk] or
        'corr' [nell, ns] (the P -> xi FFTLog operator applied at plan time or by the FFTLog kernel)."""
        from .plan import Plan
        tspec = dict(theory.theory_spec(), params=list(names))
        kind = tspec['kind']
        if kind == 'kaiser' and tspec.get('s') is not None:
            # correlation-function Kaiser theories: what is emulated is ``theory.pt`` = the P(k) tables on kin (reference full_shape.py:341-346); the
            # P -> xi operator is folded into the emulator's output maps by the tracer class (theories.KaiserTracerCorrelationFunctionMultipoles)
            tspec = {key: value for key, value in tspec.items() if key not in ('s', 'kin', 'interp_order')}
        if kind == 'kaiser':
            tables, x, cls_ = ['pk_dd', 'pk_dt', 'pk_tt'], np.asarray(theory.k, dtype='f8'), ('desilike.theories.galaxy_clustering.full_shape', 'KaiserPowerSpectrumMultipoles')
        elif kind == 'bao_pk':
            tables, x, cls_ = ['power'], np.asarray(theory.k, dtype='f8'), ('desilike.theories.galaxy_clustering.bao', 'DampedBAOWigglesTracerPowerSpectrumMultipoles')
        elif kind == 'bao_xi':
            tables, x, cls_ = ['corr'], np.asarray(theory.s, dtype='f8'), ('desilike.theories.galaxy_clustering.bao', 'DampedBAOWigglesTracerCorrelationFunctionMultipoles')
        else:
            raise ValueError('emulator construction expects a direct Kaiser or damped-BAO tracer theory (got kind {!r})'.format(kind))
        nin = len(x) * len(theory.ells)
        obs = dict(name='fit', theory=tspec, terms=[dict(scale=1., factors=[], table=tables[0])], nin=nin, wmatrix=None, offset=None, data=np.zeros(nin))
        plan = Plan(dict(params=list(names), priors={}, solved=[], observables=[obs], precision=[np.ones(nin)]), with_theory=False, merge_terms=False)
        gen = plan.lowered.gens[0]

        def evaluate(X):
            feats = plan.features(0, np.ascontiguousarray(X, dtype='f8'))
            out = {}
            for name in tables:
                foff, K, D = gen.tables[name]
                out[name] = feats[:, foff:foff + K] if D is None else feats[:, foff:foff + K] @ D
            return out

        evaluate.close = plan.close
        fixed = {('s' if kind == 'bao_xi' else 'k'): x, 'ells': tuple(theory.ells), 'z': getattr(theory.template, 'z', 1.)}
        return evaluate, {name: (len(theory.ells), len(x)) for name in tables}, fixed, cls_

    @classmethod
    def fit_taylor(cls, theory, order=4, delta=None, params=None, nbatch=None):
        """
        Build a Taylor emulator of a direct tracer theory's tables -- Kaiser: the perturbation-theory tables (``KaiserTracerPowerSpectrumMultipoles``);
        damped BAO: the P_ell(k) / xi_ell(s) multipoles without broadband -- the ``.fit`` half of the reference's emulator API
        (``Emulator(calculator, engine=TaylorEmulatorEngine(order=4)).set_samples().fit()``, reference emulators/__init__.py:438-507;
        docs/user/getting_started.rst:281-288).

        The reference differentiates the calculator with ``Differentiation`` (jax autodiff or finite differences, differentiation.py:469-551)
        and stores ``derivatives[t] = d^alpha y / alpha!`` for the term list ``powers`` (emulators/__init__.py:471-507).  Here the theory
        tables are evaluated ON THE GPU (the theory kernels, through ``dlb_features_batch``) on the tensor-product stencil of
        ``order + 1`` (odd) nodes per parameter around the parameters' default values, and the Taylor coefficients follow from the
        exact polynomial interpolation of the stencil (central finite differences, exact for polynomials of degree <= order).

        ``delta``: stencil spacing per parameter (default: ``Parameter.delta`` or 0.02); ``params``: subset of the theory's varied
        parameters entering the tables (default: all of them).
        """
        import itertools
        tspec = theory.theory_spec()
        names = list(params) if params is not None else [n for n in tspec['params'] if not theory.all_params[n].fixed]
        ndim = len(names)
        center = np.array([float(theory.all_params[n].value) for n in names])
        if delta is None:
            delta = [getattr(theory.all_params[n], 'delta', None) or 0.02 for n in names]
        delta = np.broadcast_to(np.asarray(delta, dtype='f8'), (ndim,)).copy()
        # term list in the enumeration order of the reference's fit (emulators/__init__.py:485-503)
        powers = []
        for o in range(order + 1):
            for indices in itertools.product(range(ndim), repeat=o):
                orders = tuple(int(c) for c in np.bincount(indices, minlength=ndim)) if o else (0,) * ndim
                if orders not in powers: powers.append(orders)
        powers = np.array(powers, dtype='i4')
        npts = order + 1 if (order + 1) % 2 else order + 2
        nodes = np.arange(npts, dtype='f8') - npts // 2
        grids = np.meshgrid(*[nodes] * ndim, indexing='ij')
        X = np.stack([g.ravel() for g in grids], axis=-1) * delta + center
        evaluate, yshapes, fixed, calculator_class = cls._theory_tables(theory, names)
        values = evaluate(X)
        evaluate.close()
        tables = {}
        Vinv = np.linalg.inv(np.vander(nodes, npts, increasing=True))
        for name, yshape in yshapes.items():
            Y = values[name].reshape((npts,) * ndim + (-1,))
            for ax in range(ndim):
                Y = np.moveaxis(np.tensordot(Vinv, np.moveaxis(Y, ax, 0), axes=(1, 0)), 0, ax)
            tables[name] = np.array([Y[tuple(p)] / np.prod(delta**p) for p in powers]).reshape((len(powers),) + tuple(yshape))
        all_params = [theory.all_params[n] for n in tspec['params'] if n in theory.all_params.names()] if hasattr(theory.all_params, 'names') else None
        return cls.from_taylor(names, center, powers, tables, fixed, all_params=all_params, calculator_class=calculator_class)

    @staticmethod
    def qmc_points(lower, upper, niterations, engine='rqrs', seed=0.5, offset=0):
        """Quasi-Monte-Carlo design of the reference's ``QMCSampler`` (desilike/samplers/qmc.py:13-36, :122-126): the R additive recurrence
        ('rqrs', the default: point n = (seed + n phi^-(1 + j)) mod 1 per dimension j, phi the root of phi^(d+1) = phi + 1) or one of SciPy's
        engines ('sobol', 'halton', 'lhs'), scaled to [lower, upper]."""
        lower, upper = np.asarray(lower, dtype='f8'), np.asarray(upper, dtype='f8')
        d = lower.size
        if engine == 'rqrs':
            phi = 1.0
            check = phi**(d + 1) - phi - 1
            while abs(check) > 1e-12:
                phi -= (phi**(d + 1) - phi - 1) / ((d + 1) * phi**d - 1)
                check = phi**(d + 1) - phi - 1
            inv_phi = np.array([phi**(-(1 + j)) for j in range(d)])
            u = (float(seed) + np.arange(offset + 1, offset + niterations + 1)[:, None] * inv_phi) % 1.
        else:
            from scipy.stats import qmc
            eng = {'sobol': qmc.Sobol, 'halton': qmc.Halton, 'lhs': qmc.LatinHypercube}[engine](d=d, seed=None if seed == 0.5 else int(seed))
            if offset: eng.fast_forward(offset)
            u = eng.random(n=niterations)
        return lower + u * (upper - lower)

    @classmethod
    def qmc_samples(cls, theory, niterations=int(1e5), engine='rqrs', params=None, seed=0.5, offset=0, nbatch=1 << 18):
        """
        Training samples of an MLP emulator (reference ``MLPEmulatorEngine.get_default_samples``, emulators/__init__.py:510-533: ``QMCSampler``
        with the 'rqrs' engine, 1e5 points in the box value +- proposal of every varied parameter, samplers/qmc.py:112-126), the theory tables
        evaluated ON THE GPU in batches through the theory kernels.  -> dict(params=[names], X [n, d], Y={table: [n, *yshape]}, lower, upper).
        The network training itself is not part of this module.
        """
        tspec = theory.theory_spec()
        names = list(params) if params is not None else [n for n in tspec['params'] if not theory.all_params[n].fixed]
        lower, upper = [], []
        for n in names:
            p = theory.all_params[n]
            prop = getattr(p, 'proposal', None)
            if prop is None:   # reference Parameter.proposal: scale of the ref (else prior) distribution
                for dist in (p.ref, p.prior):
                    if dist is None: continue
                    if dist.get('dist', 'uniform') == 'uniform' and dist.get('limits') is not None and np.all(np.isfinite(dist['limits'])):
                        prop = 0.5 * (dist['limits'][1] - dist['limits'][0]); break
                    if 'scale' in dist: prop = float(dist['scale']); break
            if prop is None: raise ValueError('Provide parameter limits or proposal for {}'.format(n))
            lower.append(p.value - prop); upper.append(p.value + prop)
        X = cls.qmc_points(lower, upper, niterations, engine=engine, seed=seed, offset=offset)
        evaluate, yshapes, fixed, _ = cls._theory_tables(theory, names)
        Y = {name: np.empty((niterations,) + tuple(shape)) for name, shape in yshapes.items()}
        for start in range(0, niterations, nbatch):
            vals = evaluate(X[start:start + nbatch])
            for name in Y: Y[name][start:start + nbatch] = vals[name].reshape((-1,) + tuple(yshapes[name]))
        evaluate.close()
        return dict(params=names, X=X, Y=Y, lower=np.array(lower), upper=np.array(upper), fixed=fixed)

    @classmethod
    def from_mlp(cls, params, mlp, tables, yshapes, fixed, defaults=None, all_params=None,
                 calculator_class=('desilike.theories.galaxy_clustering.full_shape', 'KaiserPowerSpectrumMultipoles')):
        """``mlp``: dict(xlimits, layers [(kernel, bias)], activation, ylimits, pca); ``tables``: {name: slice into the network output}."""
        defaults = defaults or {p: 0. for p in params}
        state = cls._base_state(params, defaults, fixed, list(tables), calculator_class, all_params)
        xops = [_op(_MINMAX, locals={'limits': np.asarray(mlp['xlimits'])})] if mlp.get('xlimits') is not None else []
        hidden, (W, b) = mlp['layers'][:-1], mlp['layers'][-1]
        act = _ACT_EXPR[mlp.get('activation', 'silu')]
        for name, sl in tables.items():
            ops = []
            for kernel, bias in hidden:
                ops += [_op(_DENSE, locals={'kernel': np.asarray(kernel), 'bias': np.asarray(bias)}), _op(act)]
            if mlp.get('pca') is not None:
                raise NotImplementedError('writer: per-table PCA')
            ops.append(_op(_DENSE, locals={'kernel': np.asarray(W)[:, sl], 'bias': np.asarray(b)[sl]}))
            yops = [_op('(' + _MINMAX + ')', inverse=_MINMAX_INV, locals={'limits': np.asarray(mlp['ylimits'])[sl]})] if mlp.get('ylimits') is not None else []
            state['engines'][name] = {'name': 'mlp', 'params': list(params), 'xshape': (len(params),), 'yshape': tuple(yshapes[name]),
                                      'xoperations': xops, 'yoperations': yops, 'model_operations': ops, 'model_yoperations': []}
        return cls(state)


def _parse_mlp_engine(engine):
    """-> dict(xlimits, layers, activation, ylimits, pca) of ONE engine, by pattern-matching its operation strings."""
    out = dict(xlimits=None, ylimits=None, pca=None, activation='identity', layers=[])
    for op in engine.get('xoperations', []):
        expr, inv, loc = _op_fields(op)
        if _norm(expr) != _norm(_MINMAX): raise ValueError('x-operation {!r} is not on the CUDA path'.format(expr))
        out['xlimits'] = np.asarray(loc['limits'], dtype='f8')
    acts = set()
    for op in engine['model_operations']:
        expr, inv, loc = _op_fields(op)
        e = _norm(expr)
        if e == _norm(_DENSE):
            out['layers'].append((np.asarray(loc['kernel'], dtype='f8'), np.asarray(loc['bias'], dtype='f8')))
        elif e in {_norm(k): v for k, v in _ACTS.items()}:
            acts.add({_norm(k): v for k, v in _ACTS.items()}[e])
        else:
            raise ValueError('model operation {!r} is not on the CUDA path (dense / silu / relu / tanh only)'.format(expr))
    if len(acts) > 1: raise ValueError('mixed activations are not on the CUDA path')
    if acts: out['activation'] = acts.pop()
    if engine.get('model_yoperations'): raise ValueError('model_yoperations are not on the CUDA path')
    ny = out['layers'][-1][0].shape[1]
    for op in engine.get('yoperations', []):   # inverse operations are applied in reverse order at predict time
        expr, inv, loc = _op_fields(op)
        if _norm(inv) == _norm(_MINMAX_INV):
            out['ylimits'] = np.asarray(loc['limits'], dtype='f8').reshape(ny, 2)
        elif _norm(inv) == _norm(_PCA):
            out['pca'] = (np.asarray(loc['eigvecs'], dtype='f8'), np.asarray(loc['mean'], dtype='f8'))
        else:
            raise ValueError('y-operation with inverse {!r} is not on the CUDA path (min-max scaling / PCA only)'.format(inv))
    if len(engine.get('yoperations', [])) > 1: raise ValueError('chained y-operations are not on the CUDA path')
    return out


def _merge_mlps(nets):
    """Merge per-table networks into ONE network whose output is the concatenation of theirs: shared hidden stacks are kept
    once, different ones are stacked block-diagonally (exact: the extra entries are zeros)."""
    names = list(nets)
    first = nets[names[0]]
    same_hidden = all(len(n['layers']) == len(first['layers']) and all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                      for a, b in zip(n['layers'][:-1], first['layers'][:-1])) for n in nets.values())
    for n in nets.values():
        if n['activation'] != first['activation'] or (n['xlimits'] is None) != (first['xlimits'] is None) or \
           (n['xlimits'] is not None and not np.array_equal(n['xlimits'], first['xlimits'])) or n['pca'] is not None:
            raise ValueError('engines with different x-operations / activations (or PCA) cannot be merged on the CUDA path')
    sizes = [n['layers'][-1][0].shape[1] for n in nets.values()]
    tables, start = {}, 0
    for name, size in zip(names, sizes):
        tables[name] = slice(start, start + size); start += size
    if same_hidden:
        layers = list(first['layers'][:-1])
        layers.append((np.concatenate([n['layers'][-1][0] for n in nets.values()], axis=1), np.concatenate([n['layers'][-1][1] for n in nets.values()])))
    else:
        import scipy.linalg
        nl = len(first['layers'])
        if any(len(n['layers']) != nl for n in nets.values()): raise ValueError('engines with different depths cannot be merged')
        layers = [(np.concatenate([n['layers'][0][0] for n in nets.values()], axis=1), np.concatenate([n['layers'][0][1] for n in nets.values()]))]
        for il in range(1, nl):
            layers.append((scipy.linalg.block_diag(*[n['layers'][il][0] for n in nets.values()]), np.concatenate([n['layers'][il][1] for n in nets.values()])))
        if nl == 1: layers = [(np.concatenate([n['layers'][0][0] for n in nets.values()], axis=1), np.concatenate([n['layers'][0][1] for n in nets.values()]))]
    ylimits = None
    if any(n['ylimits'] is not None for n in nets.values()):
        ylimits = np.concatenate([n['ylimits'] if n['ylimits'] is not None else np.stack([np.zeros(s), np.ones(s)], axis=-1) for n, s in zip(nets.values(), sizes)])
    return dict(xlimits=first['xlimits'], layers=layers, activation=first['activation'], ylimits=ylimits, pca=None), tables


class EmulatedCalculator(BaseCalculator):
    """Emulated perturbation-theory calculator (reference emulators/__init__.py:394-418): pass it as ``pt=`` to a tracer theory."""

    def __init__(self, emulator):
        self.emulator = emulator
        state = emulator.state
        self.fixed = state['fixed']
        self.k, self.ells = np.asarray(self.fixed['k'] if 'k' in self.fixed else self.fixed['s'], dtype='f8'), tuple(self.fixed['ells'])
        self.s = self.k if 's' in self.fixed else None
        self.varied_params_names = list(state.get('varied_params', []))
        params = []
        for ps in state.get('all_params', {}).get('data', []):
            params.append(Parameter(ps['basename'], value=ps.get('value', 0.), fixed=ps.get('fixed', False), prior=ps.get('prior'), ref=ps.get('ref'),
                                    derived=ps.get('derived', False), namespace=ps.get('namespace', '')))
        self.params = ParameterCollection(params)

    @classmethod
    def load(cls, filename):
        return Emulator.load(filename).to_calculator()

    def save(self, filename):
        return self.emulator.save(filename)

    def theory_spec(self, tracer_theory=None):
        state = self.emulator.state
        engines = state['engines']
        if tracer_theory is not None:
            if len(tracer_theory.k) != len(self.k) or not np.allclose(tracer_theory.k, self.k) or tuple(tracer_theory.ells) != self.ells:
                raise ValueError('the emulator was built on a different (k, ells) grid than the one requested by the window / observable')
        kinds = {e['name'] for e in engines.values()}
        if len(kinds) != 1: raise ValueError('mixed emulator engines are not on the CUDA path')
        kind = kinds.pop()
        nk, nell = len(self.k), len(self.ells)
        names = list(engines)
        first = engines[names[0]]
        params = list(first['params'])
        defaults = dict(state.get('defaults', {}))
        if kind == 'taylor':
            for e in engines.values():
                if list(e['params']) != params or not np.array_equal(e['center'], first['center']) or not np.array_equal(e['powers'], first['powers']):
                    raise ValueError('Taylor engines with different expansion points / term lists are not on the CUDA path')
                if e.get('xoperations') or e.get('yoperations'): raise ValueError('operations are not supported for the Taylor engine (reference emulators/__init__.py:435-436)')
            T = np.asarray(first['powers']).shape[0]
            tables = {name: np.asarray(e['derivatives'], dtype='f8').reshape(T, -1) for name, e in engines.items()}
            src = tables.get('pk11', tables.get('pk_dd'))
            if src is not None and 0 in self.ells:
                i0 = self.ells.index(0)
                tables['pk11_l0'] = np.tile(src.reshape(T, nell, nk)[:, i0], (1, nell))
            return dict(kind='taylor', params=params, center=np.asarray(first['center'], dtype='f8'), powers=np.asarray(first['powers'], dtype='i4'),
                        tables=tables, defaults=defaults)
        if kind == 'mlp':
            nets = {name: _parse_mlp_engine(e) for name, e in engines.items()}
            mlp, tables = _merge_mlps(nets)
            src = 'pk11' if 'pk11' in tables else 'pk_dd' if 'pk_dd' in tables else None
            if src is not None and 0 in self.ells:
                i0 = self.ells.index(0)
                base = np.arange(tables[src].start, tables[src].stop).reshape(nell, nk)[i0]
                tables['pk11_l0'] = np.tile(base, nell)
            return dict(kind='mlp', params=params, mlp=mlp, tables=tables, defaults=defaults)
        raise ValueError('emulator engine {!r} is not on the CUDA path'.format(kind))

    def scalar_expression(self, name):
        """A scalar output of the emulated calculator (e.g. ``sigma8``, ``fsigma8`` of a perturbation-theory calculator) as a derived-parameter
        EXPRESSION of the emulator's parameters (desilike_b200/expressions.py): a constant when the state holds it as a fixed entry, the Taylor
        polynomial sum_t D_t prod_j ({p_j} - c_j)^pow_tj for a scalar Taylor engine (layout of reference emulators/__init__.py:471-507)."""
        state = self.emulator.state
        if name in state['engines']:
            e = state['engines'][name]
            if e['name'] != 'taylor' or int(np.prod(e['yshape'])) != 1:
                raise ValueError('scalar {!r}: only scalar Taylor engines can feed derived parameters on the CUDA path (got {!r}, yshape {})'.format(name, e['name'], e['yshape']))
            center, powers, deriv = np.asarray(e['center'], dtype='f8'), np.asarray(e['powers'], dtype='i4'), np.asarray(e['derivatives'], dtype='f8').ravel()
            out = []
            for t in range(len(powers)):
                if deriv[t] == 0.: continue
                fac = ['({!r})'.format(float(deriv[t]))]
                for j, p in enumerate(e['params']):
                    if powers[t, j]: fac.append('({{{}}} - ({!r}))**{:d}'.format(p, float(center[j]), int(powers[t, j])))
                out.append(' * '.join(fac))
            return ' + '.join(out) if out else '0.'
        if name in state['fixed']: return repr(float(state['fixed'][name]))
        raise ValueError('the emulator provides no {!r} (neither an engine nor a fixed entry)'.format(name))

    def table_spec(self, name):
        """Taylor engine ``name`` as (theory spec without tables, derivatives [T, *yshape]) -- for tracer theories that post-process the
        emulated table themselves (``LPTVelocileptorsTracerPowerSpectrumMultipoles``: bias combination + re-interpolation in k)."""
        state = self.emulator.state
        if name not in state['engines']: raise ValueError('emulator has no engine {!r} (engines: {})'.format(name, sorted(state['engines'])))
        e = state['engines'][name]
        if e['name'] == 'mlp':   # -> (spec, yshape): the table is a slice of the network output, selected by index
            spec = dict(kind='mlp', params=list(e['params']), mlp=_parse_mlp_engine(e), tables={}, defaults=dict(state.get('defaults', {})))
            return spec, tuple(e['yshape'])
        if e['name'] != 'taylor': raise ValueError('table models on the CUDA path need a Taylor or MLP emulator of {!r} (got {!r})'.format(name, e['name']))
        if e.get('xoperations') or e.get('yoperations'): raise ValueError('operations are not supported for the Taylor engine (reference emulators/__init__.py:435-436)')
        deriv = np.asarray(e['derivatives'], dtype='f8')
        spec = dict(kind='taylor', params=list(e['params']), center=np.asarray(e['center'], dtype='f8'), powers=np.asarray(e['powers'], dtype='i4'),
                    tables={}, defaults=dict(state.get('defaults', {})))
        return spec, deriv
