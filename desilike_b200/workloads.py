"""Seeded synthetic workloads for BASELINE.json configs C1-C5, as likelihood specs.

Product-side generators of synthetic constants (template grids, window matrices, covariances, emulator coefficients): what ``bench.py`` builds its
inputs from, and -- re-exported by ``oracle/configs.py`` -- what the parity tests evaluate through BOTH the oracle and the CUDA path.  No reference
arithmetic lives here.  Two steps need a theory evaluation: the synthetic data vector (theory at a fiducial point) and the optional finite-difference
Taylor coefficients of configuration C2(i).  They go through hooks: by default the package's own CUDA path (``Plan.theory``), so that ``bench.py`` builds
its workloads without importing the checker; ``oracle/configs.py`` installs the oracle's functions instead (CPU test suite, reference arm).
Shapes follow SURVEY.md section 8(d); sizes can be shrunk for fast tests.
"""
import re
import itertools

import numpy as np

F_FID = 0.8

_hooks = dict(theory_flat=None, theory_tables=None)


def set_hooks(theory_flat=None, theory_tables=None):
    """``theory_flat(obs, points, names, x0) -> [N, n]`` flat theory of one observable at the solved defaults ``x0``; ``theory_tables(theory, points, names)
    -> dict of pt tables`` (only for ``coefficients='finite-difference'``).  ``None`` restores the default (the CUDA path)."""
    _hooks['theory_flat'], _hooks['theory_tables'] = theory_flat, theory_tables


def _observable_flat(obs, points, names, x0):
    if _hooks['theory_flat'] is not None: return _hooks['theory_flat'](obs, points, names, x0)
    # default: the product itself -- a one-observable plan with a placeholder data vector; theory vectors do not depend on data or precision
    from .plan import Plan
    n = obs['wmatrix'].shape[0] if obs.get('wmatrix') is not None else int(obs['nin'])
    spec = dict(params=list(names), observables=[dict(obs, data=np.ones(n), transform=None)], precision=[np.ones(n)], priors=default_priors(names),
                solved=[dict(name=name, value=value, prior=None, marg=True) for name, value in x0.items()])
    plan = Plan(spec, reduce_rank=False, qform=False)
    try: return plan.theory(np.asarray(points, dtype=np.float64))
    finally: plan.close()


def _theory_tables(theory, points, names):
    if _hooks['theory_tables'] is None:
        raise RuntimeError("coefficients='finite-difference' needs a theory_tables hook (oracle/configs.py installs the oracle's); use coefficients='random'")
    return _hooks['theory_tables'](theory, points, names)


def pknow_fid(k):
    """Smooth BBKS-like broadband, amplitude ~2e4 (Mpc/h)^3 at k = 0.02 h/Mpc (closed-form stand-in for the cosmoprimo fiducial)."""
    k = np.asarray(k, dtype='f8')

    def shape(kk):
        q = kk / 0.2 * 1.45
        T = np.log(1. + 2.34 * q) / (2.34 * q) * (1. + 3.89 * q + (16.1 * q)**2 + (5.46 * q)**3 + (6.71 * q)**4)**(-0.25)
        return kk**0.965 * T**2
    return 2e4 * shape(k) / shape(0.02)


def pk_fid(k):
    """Broadband times a damped BAO-like wiggle."""
    k = np.asarray(k, dtype='f8')
    return pknow_fid(k) * (1. + 0.05 * np.sin(k * 105.) * np.exp(-(k * 8.)**2))


def eftlike_matrices(k, ells, ct_names, sn_names, kp=1.):
    """Counter-term / stochastic templates [nell, nk, nparams] for names 'ct{ell}_{pow}', 'sn{ell}_{pow}', 'sn0' (ell 0, constant): (k / kp)^pow in the
    multipole the name points at (layout of reference full_shape.py:590-621)."""
    def matrix(names, base):
        mat = np.zeros((len(ells), len(k), len(names)))
        for i, name in enumerate(names):
            if name == base + '0': ell, col = 0, np.ones_like(k)
            else:
                ell, power = [int(v) for v in re.match(base + '(.*)_(.*)', name).groups()]
                col = (k / kp)**power
            if ell in ells: mat[list(ells).index(ell), :, i] = col
        return mat
    return matrix(ct_names, 'ct'), matrix(sn_names, 'sn')


def taylor_terms(ndim, order):
    """Multi-indices of all monomials up to ``order`` in the enumeration order of reference emulators/__init__.py:485-503."""
    powers = []
    for o in range(order + 1):
        for indices in itertools.product(range(ndim), repeat=o):
            p = tuple(np.bincount(indices, minlength=ndim).astype('i4')) if o else (0,) * ndim
            if p not in powers: powers.append(p)
    return np.array(powers, dtype='i4')


def taylor_fit_polynomial(func, center, ndim, order, delta):
    """Taylor coefficients d^alpha y / alpha! at ``center`` from a tensor-product stencil (exact for polynomials of degree <= order)."""
    powers = taylor_terms(ndim, order)
    npts = order + 1 + (order + 1) % 2 * 0
    if npts % 2 == 0: npts += 1
    nodes = (np.arange(npts) - npts // 2).astype('f8')
    Vinv = np.linalg.inv(np.vander(nodes, npts, increasing=True))
    grids = np.meshgrid(*[nodes] * ndim, indexing='ij')
    Y = func(np.stack([g.ravel() for g in grids], axis=-1) * delta + center)
    Y = Y.reshape((npts,) * ndim + Y.shape[1:])
    for ax in range(ndim): Y = np.moveaxis(np.tensordot(Vinv, np.moveaxis(Y, ax, 0), axes=(1, 0)), 0, ax)
    return powers, np.array([Y[tuple(p)] / np.prod(np.asarray(delta, dtype='f8')**p) for p in powers])


def klim_edges(klim=(0.02, 0.2, 0.005)):
    edges = np.arange(klim[0], klim[1] + 1e-9, klim[2])
    return (edges[:-1] + edges[1:]) / 2.


def shapefit_template(nk=500, klim=(1e-3, 1.)):
    """Constants of ShapeFitPowerSpectrumTemplate (reference power_template.py:747-761) on the PT grid
    (reference full_shape.py:19,29: geomspace(1e-3, 1, 500))."""
    k = np.geomspace(klim[0], klim[1], nk)
    return dict(kind='shapefit', k=k, pk_dd_fid=pk_fid(k), pknow_dd_fid=pknow_fid(k), f_fid=F_FID,
                a=0.6, kp=0.03, apmode='qparqper', eta=1. / 3.)


def bao_template(nk=2000, klim=(1e-4, 1.)):
    """Constants of BAOPowerSpectrumTemplate on the BAO grid (reference bao.py:67,79: geomspace(1e-4, 1, 2000))."""
    k = np.geomspace(klim[0], klim[1], nk)
    return dict(kind='bao', k=k, pk_dd_fid=pk_fid(k), pknow_dd_fid=pknow_fid(k), f_fid=F_FID,
                apmode='qparqper', eta=1. / 3.)


def gaussian_covariance(k, ells, pk0, volume=1e10, dk=0.005, coupling=0.1, rng=None):
    """Gaussian-diagonal 2 (2 pi)^3 / ((2 ell + 1) V N_modes) P0^2 + ``coupling`` random SPD mixing."""
    blocks = []
    for ell in ells:
        nmodes = 4. * np.pi * k**2 * dk
        blocks.append(2. * (2. * np.pi)**3 / volume / nmodes * (2 * ell + 1) * pk0**2)
    sigma = np.sqrt(np.concatenate(blocks))
    n = sigma.size
    corr = np.eye(n)
    if coupling and rng is not None:
        A = rng.standard_normal((n, max(n // 4, 2)))
        C = A.dot(A.T)
        C /= np.sqrt(np.outer(np.diag(C), np.diag(C)))
        corr = (1. - coupling) * np.eye(n) + coupling * C
    return corr * np.outer(sigma, sigma)


def window_matrix(kout, ells, kin, ellsin, width=0.004, leakage=0.05):
    """Banded Gaussian rows + ``leakage`` into the other multipoles, rows normalised."""
    rows = []
    for ello in ells:
        for ko in kout:
            row = []
            for elli in ellsin:
                w = np.exp(-0.5 * ((kin - ko) / width)**2)
                w /= w.sum()
                row.append(w * (1. if elli == ello else leakage))
            rows.append(np.concatenate(row))
    return np.array(rows)


def kaiser_theory(k, ells=(0, 2, 4), nmu=8, template=None, params=('qpar', 'qper', 'dm', 'df')):
    template = template or shapefit_template()
    return dict(kind='kaiser', k=np.asarray(k, dtype='f8'), ells=tuple(ells), nmu=nmu, template=template, params=list(params),
                defaults=dict(qpar=1., qper=1., dm=0., dn=0., df=1., sigmapar=0., sigmaper=0.))


def kaiser_terms(ells, nk, nd, b1='b1', sn0='sn0'):
    """Tracer combine, auto-correlation (reference full_shape.py:545-550)."""
    e0 = np.concatenate([np.full(nk, float(ell == 0)) for ell in ells])
    return [dict(scale=1., factors=[(b1, 2)], table='pk_dd'), dict(scale=2., factors=[(b1, 1)], table='pk_dt'),
            dict(scale=1., factors=[], table='pk_tt'), dict(scale=1. / nd, factors=[(sn0, 1)], table=None, colvec=e0)]


def eftlike_terms(k, ells, nd, ct_names, sn_names, prefix=''):
    """EFT-like counter / stochastic add-ons as *solved* linear terms (reference full_shape.py:586-634)."""
    ctm, snm = eftlike_matrices(k, ells, ct_names, sn_names, kp=1.)
    terms = []
    for i, name in enumerate(ct_names):  # auto-correlation: 0.5 * (ct + ct) = ct
        terms.append(dict(scale=1., factors=[], solved=prefix + name, table='pk11_l0', colvec=ctm[:, :, i].ravel()))
    for i, name in enumerate(sn_names):
        # reference behaviour: in the EFT-like class 'sn0' enters TWICE -- once through KaiserTracer.calculate
        # (full_shape.py:549) and once through the stochastic matrix (full_shape.py:634) -- hence the factor 2
        scale = (2. if name == 'sn0' else 1.) / nd
        terms.append(dict(scale=scale, factors=[], solved=prefix + name, table=None, colvec=snm[:, :, i].ravel()))
    return terms


_POINT_RANGES = dict(qpar=('u', 0.95, 1.05), qper=('u', 0.95, 1.05), dm=('u', -0.05, 0.05), df=('u', 0.9, 1.1), b1=('u', 1.5, 2.5),
                     sn0=('n', 0., 0.1), dbeta=('u', 0.8, 1.2), sigmas=('u', 2., 8.), sigmapar=('u', 2., 8.), sigmaper=('u', 2., 8.))


def sample_points(spec, N, seed=0):
    """Points as in SURVEY.md section 8d ('points as C1')."""
    rng = np.random.default_rng(seed)
    X = np.empty((N, len(spec['params'])))
    for i, name in enumerate(spec['params']):
        base = name.split('.')[-1]
        kind, a, b = _POINT_RANGES.get(base, ('n', 0., 1.))
        X[:, i] = rng.uniform(a, b, N) if kind == 'u' else rng.normal(a, b, N)
    return X


def default_priors(names):
    priors = {}
    for name in names:
        base = name.split('.')[-1]
        if base in ('qpar', 'qper'): priors[name] = dict(dist='uniform', limits=(0.8, 1.2))
        elif base == 'dm': priors[name] = dict(dist='uniform', limits=(-3., 3.))
        elif base == 'df': priors[name] = dict(dist='uniform', limits=(0., 2.))
        elif base == 'b1': priors[name] = dict(dist='uniform', limits=(0., 10.))
        elif base == 'sn0': priors[name] = dict(dist='norm', loc=0., scale=2., limits=(-10., 10.))
        elif base == 'dbeta': priors[name] = dict(dist='uniform', limits=(0.25, 1.75))
        elif base == 'sigmas': priors[name] = dict(dist='norm', loc=2., scale=2., limits=(0., 20.))
        elif base in ('sigmapar', 'sigmaper'): priors[name] = dict(dist='norm', loc=6., scale=3., limits=(0., 20.))
        else: priors[name] = None
    return priors


def config_c1(seed=1):
    """C1: Kaiser + ShapeFit -> (no window) -> chi2; ells (0,2,4), 36 bins/ell, n = 108, d = 6."""
    rng = np.random.default_rng(seed)
    ells = (0, 2, 4)
    k = klim_edges()
    nd = 1e-4
    theory = kaiser_theory(k, ells)
    obs = dict(name='LRG', theory=theory, terms=kaiser_terms(ells, k.size, nd), nin=len(ells) * k.size, wmatrix=None, offset=None)
    spec = dict(params=['qpar', 'qper', 'dm', 'df', 'b1', 'sn0'], observables=[obs], solved=[])
    spec['priors'] = default_priors(spec['params'])
    fid = np.array([[1., 1., 0., 1., 2., 0.]])
    obs['data'] = _observable_flat(obs, fid, spec['params'], {})[0]
    cov = gaussian_covariance(k, ells, obs['data'][:k.size] + 1. / nd, rng=rng)
    spec['precision'] = [np.linalg.inv(cov)]
    return spec


def _kaiser_taylor_engine(kin, ells, order, template, delta=(0.02, 0.02, 0.02, 0.04), with_pk11=False):
    """Taylor coefficients (layout of reference emulators/__init__.py:471-507) of the Kaiser pt tables in (qpar,qper,dm,df)
    by finite differences -- C2 coefficient option (i) of SURVEY.md section 8d."""
    theory = kaiser_theory(kin, ells, template=template)
    names = theory['params']
    tabnames = ['pk_dd', 'pk_dt', 'pk_tt'] + (['pk11_l0'] if with_pk11 else [])

    def func(pts):
        tabs = _theory_tables(theory, pts, names)
        return np.concatenate([tabs[n] for n in tabnames], axis=-1)

    center = np.array([1., 1., 0., 1.])
    powers, derivs = taylor_fit_polynomial(func, center, 4, order, np.asarray(delta))
    nin = len(ells) * len(kin)
    tables = {n: derivs[:, i * nin:(i + 1) * nin].copy() for i, n in enumerate(tabnames)}
    return dict(kind='taylor', params=list(names), center=center, powers=powers, tables=tables)


def _random_taylor_engine(kin, ells, order, rng, with_pk11=False, ndim=4):
    """C2 coefficient option (ii): N(0,1) * 0.1^order(alpha) * P_broadband(k) (throughput-shaped, no Kaiser evaluations)."""
    powers = taylor_terms(ndim, order)
    nin = len(ells) * len(kin)
    amp = np.tile(pknow_fid(kin), len(ells))
    scale = 0.1**powers.sum(axis=-1)
    tables = {}
    for name in ['pk_dd', 'pk_dt', 'pk_tt']:
        d = rng.standard_normal((len(powers), nin)) * scale[:, None] * amp
        d[0] = amp * {'pk_dd': 1., 'pk_dt': 0.5, 'pk_tt': 0.3}[name]
        tables[name] = d
    if with_pk11:
        nk = len(kin)
        tables['pk11_l0'] = np.tile(tables['pk_dd'][:, :nk], (1, len(ells)))
    return dict(kind='taylor', params=['qpar', 'qper', 'dm', 'df'], center=np.array([1., 1., 0., 1.]), powers=powers, tables=tables)


def taylor_tracer(name='', seed=2, order=4, nkin=350, nbins=36, coefficients='finite-difference', ells=(0, 2, 4), shared_template=True,
                  solved=False, rng=None, template=None):
    """One Taylor-emulated tracer: pt tables [3, nkin] x3 -> Kaiser combine -> window [3*nbins, 3*nkin]."""
    rng = rng or np.random.default_rng(seed)
    prefix = (name + '.') if name else ''
    kin = np.arange(1, nkin + 1) * (0.35 / nkin)
    # bin width 0.005 (reference notebooks) unless that would run past the emulator's k range (kin <= 0.35): then the bins
    # share [0.02, 0.34] (5-tracer configs with 134 bins per multipole)
    dk = min(0.005, (0.34 - 0.02) / nbins)
    kout = 0.02 + dk * (np.arange(nbins) + 0.5)
    nd = 1e-4
    if coefficients == 'finite-difference':
        engine = _kaiser_taylor_engine(kin, ells, order, template or shapefit_template(), with_pk11=solved)
    else:
        engine = _random_taylor_engine(kin, ells, order, rng, with_pk11=solved)
    emu_params = ['qpar', 'qper', 'dm', 'df'] if shared_template else [prefix + p for p in ['qpar', 'qper', 'dm', 'df']]
    engine['params'] = emu_params
    terms = kaiser_terms(ells, nkin, nd, b1=prefix + 'b1', sn0=prefix + 'sn0')
    params = list(emu_params) + [prefix + 'b1']
    solved_list = []
    if solved:
        terms = terms[:3]  # sn0 becomes a solved stochastic term
        ct_names, sn_names = ['ct0_2', 'ct2_2', 'ct4_2'], ['sn0', 'sn0_2', 'sn2_2', 'sn4_2']
        terms += eftlike_terms(kin, ells, nd, ct_names, sn_names, prefix=prefix)
        for n in ct_names: solved_list.append(dict(name=prefix + n, value=0., prior=dict(dist='norm', loc=0., scale=100.), marg=True))
        for n in sn_names: solved_list.append(dict(name=prefix + n, value=0., prior=dict(dist='norm', loc=0., scale=1000.), marg=True))
    else:
        params.append(prefix + 'sn0')
    W = window_matrix(kout, ells, kin, ells)
    shotnoise = 1. / nd
    shotnoisein = np.array([shotnoise * (ell == 0) for ell in ells])
    shotnoiseout = shotnoise * np.concatenate([np.full(nbins, float(ell == 0)) for ell in ells])
    offset = W.dot(np.repeat(shotnoisein, nkin)) - shotnoiseout  # reference window.py:445-473 folded into one constant
    obs = dict(name=name or 'tracer', theory=engine, terms=terms, nin=len(ells) * nkin, wmatrix=W, offset=offset)
    return obs, params, solved_list, (kout, ells, nd, dk)


def _finish(spec, observables_meta, rng, fid_values, coupling=0.1):
    names = spec['params']
    fid = np.array([[fid_values[n.split('.')[-1]] for n in names]])
    x0 = {s['name']: s['value'] for s in spec['solved']}
    precs = []
    for obs, meta in zip(spec['observables'], observables_meta):
        kout, ells, nd = meta[:3]
        dk = meta[3] if len(meta) > 3 else 0.005
        obs['data'] = _observable_flat(obs, fid, names, x0)[0]
        cov = gaussian_covariance(kout, ells, np.abs(obs['data'][:len(kout)]) + 1. / nd, dk=dk, rng=rng, coupling=coupling)
        precs.append(np.linalg.inv(cov))
    spec['precision'] = precs
    spec['priors'] = default_priors(names)
    return spec


_FID = dict(qpar=1., qper=1., dm=0., df=1., b1=2., sn0=0.)


def config_c2(seed=2, order=4, nkin=350, nbins=36, coefficients='finite-difference'):
    """C2: Taylor(order 4) pt tables -> Kaiser combine -> window [108, 1050] -> chi2, d = 6, T = 70."""
    rng = np.random.default_rng(seed)
    obs, params, solved, meta = taylor_tracer('', order=order, nkin=nkin, nbins=nbins, coefficients=coefficients, rng=rng)
    spec = dict(params=params, observables=[obs], solved=solved)
    return _finish(spec, [meta], rng, _FID)


def config_c5(seed=5, ntracers=5, order=4, nkin=350, nbins=134, coefficients='random', solved=False, dense=False):
    """C5: ``ntracers`` Taylor tracers sharing (qpar,qper,dm,df), own b1 (+sn0), block-diagonal precision, n ~ 5 x 402.
    ``solved=True`` gives the C4-style 7 marginalised counter/stochastic terms per tracer (with Taylor theory)."""
    rng = np.random.default_rng(seed)
    tracers = ['T{:d}'.format(i) for i in range(ntracers)]
    observables, metas, params, solved_list = [], [], ['qpar', 'qper', 'dm', 'df'], []
    for t in tracers:
        obs, p, s, meta = taylor_tracer(t, order=order, nkin=nkin, nbins=nbins, coefficients=coefficients, rng=rng, solved=solved)
        observables.append(obs); metas.append(meta); solved_list += s
        params += [q for q in p if q not in params]
    spec = dict(params=params, observables=observables, solved=solved_list)
    spec = _finish(spec, metas, rng, _FID)
    if dense: _densify(spec, rng)
    return spec


def _densify(spec, rng):
    """Joint covariance coupling the tracers (SURVEY.md section 8d C4 'dense 2000 x 2000' stress variant): block-diagonal + 2 % rank-8 coupling."""
    n = sum(len(o['data']) for o in spec['observables'])
    import scipy.linalg
    cov = scipy.linalg.block_diag(*[np.linalg.inv(p) for p in spec['precision']])
    sig = np.sqrt(np.diag(cov))
    A = rng.standard_normal((n, 8))
    C = A.dot(A.T); C /= np.sqrt(np.outer(np.diag(C), np.diag(C)))
    cov = cov + 0.02 * C * np.outer(sig, sig)
    spec['precision'] = np.linalg.inv(cov)
    return spec


def mlp_engine(rng, nin_tables, hidden=(64, 64, 64), ndim=4, activation='silu', kin=None, ells=(0, 2, 4)):
    """MLP pt-emulator: weights N(0, 2/fan_in), x min-max scale, y affine onto P_broadband scale."""
    sizes = [ndim] + list(hidden) + [nin_tables]
    layers = [(rng.standard_normal((a, b)) * np.sqrt(2. / a), 0.1 * rng.standard_normal(b)) for a, b in zip(sizes[:-1], sizes[1:])]
    xlimits = np.array([[0.9, 1.1], [0.9, 1.1], [-0.1, 0.1], [0.8, 1.2]])[:ndim]
    amp = np.tile(pknow_fid(kin), nin_tables // len(kin))
    ylimits = np.stack([0.2 * amp, 1.2 * amp], axis=-1)
    return dict(xlimits=xlimits, layers=layers, activation=activation, ylimits=ylimits, pca=None)


def config_c4(seed=4, ntracers=5, hidden=(64, 64, 64), nkin=350, nbins=134, solved=True, dense=False):
    """C4: ``ntracers`` x [MLP pt-emulator -> EFT-like Kaiser combine -> window] -> joint chi2 with analytic marginalisation."""
    rng = np.random.default_rng(seed)
    ells = (0, 2, 4)
    observables, metas, params, solved_list = [], [], [], []
    for it in range(ntracers):
        t = 'T{:d}'.format(it)
        obs, p, s, meta = taylor_tracer(t, order=1, nkin=nkin, nbins=nbins, coefficients='random', rng=rng, solved=solved, shared_template=False)
        kin = np.arange(1, nkin + 1) * (0.35 / nkin)
        nin = len(ells) * nkin
        mlp = mlp_engine(rng, 3 * nin, hidden=hidden, kin=kin)
        tables = {'pk_dd': slice(0, nin), 'pk_dt': slice(nin, 2 * nin), 'pk_tt': slice(2 * nin, 3 * nin)}
        theory = dict(kind='mlp', params=obs['theory']['params'], mlp=mlp, tables=tables)
        if solved:
            # pk11 (l=0 row of pk_dd tiled over ells): expressed as an index map into the MLP output
            idx = np.tile(np.arange(nkin), len(ells))
            theory['tables']['pk11_l0'] = idx
        obs['theory'] = theory
        observables.append(obs); metas.append(meta); solved_list += s
        params += [q for q in p if q not in params]
    spec = dict(params=params, observables=observables, solved=solved_list)
    spec = _finish(spec, metas, rng, _FID)
    if dense: _densify(spec, rng)
    return spec


def config_c3(seed=3, ns_theory=33, broadband_solved=True):
    """C3: damped-BAO P(k) -> FFTLog -> xi_ell(s) (+ broadband) -> chi2; ells (0, 2); s = arange(20, 152, 4), n = 66."""
    rng = np.random.default_rng(seed)
    ells = (0, 2)
    s = np.arange(20., 152., 4.)[:ns_theory]
    tpl = bao_template()
    theory = dict(kind='bao_xi', s=s, ells=ells, nmu=10, template=tpl, mode='', smoothing_radius=15.,
                  params=['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper'],
                  defaults=dict(qpar=1., qper=1., df=1., b1=1., dbeta=1., sigmas=0., sigmapar=9., sigmaper=6.))
    terms = [dict(scale=1., factors=[], table='corr')]
    sp = 2. * np.pi / 0.02
    solved = []
    for ill, ell in enumerate(ells):
        for power in (0, 2):  # 'even-power' broadband (reference bao.py:849-853,867,905)
            colvec = np.zeros((len(ells), s.size)); colvec[ill] = (s / sp)**power
            name = 'al{:d}_{:d}'.format(ell, power)
            terms.append(dict(scale=1., factors=[], solved=name, table=None, colvec=colvec.ravel()))
            solved.append(dict(name=name, value=0., prior=None, marg=True))
    obs = dict(name='xi', theory=theory, terms=terms, nin=len(ells) * s.size, wmatrix=None, offset=None)
    spec = dict(params=list(theory['params']), observables=[obs], solved=solved if broadband_solved else [])
    if not broadband_solved:
        obs['terms'] = terms[:1]
    fid = np.array([[1., 1., 2., 1., 4., 6., 4.]])
    obs['data'] = _observable_flat(obs, fid, spec['params'], {s_['name']: 0. for s_ in solved})[0]
    sig = 0.03 * np.abs(obs['data']) + 2e-4 * (50. / np.tile(s, len(ells)))**2
    n = sig.size
    A = rng.standard_normal((n, n // 4)); C = A.dot(A.T); C /= np.sqrt(np.outer(np.diag(C), np.diag(C)))
    cov = (0.8 * np.eye(n) + 0.2 * C) * np.outer(sig, sig)
    spec['precision'] = [np.linalg.inv(cov)]
    spec['priors'] = default_priors(spec['params'])
    return spec


def config_kaiser_window(seed=7, nkin=60, nbins=12, eftlike=False, solved=True, transform=None, nmu=8, systematic=False):
    """Direct (non-emulated) Kaiser + ShapeFit -> window -> chi2, optionally with EFT-like ct/sn terms (solved or sampled).
    Mirrors what reference classes KaiserTracerPowerSpectrumMultipoles / EFTLikeKaiserTracerPowerSpectrumMultipoles +
    WindowedPowerSpectrumMultipoles (ndarray wmatrix branch, reference window.py:309-323) compute; used to pin the oracle."""
    rng = np.random.default_rng(seed)
    ells = (0, 2, 4)
    kin = np.linspace(0.005, 0.3, nkin)
    kout = klim_edges((0.02, 0.02 + nbins * 0.01, 0.01))[:nbins]
    nd = 1e-4
    theory = kaiser_theory(kin, ells, nmu=nmu)
    params = ['qpar', 'qper', 'dm', 'df', 'b1']
    solved_list = []
    if eftlike:
        terms = kaiser_terms(ells, nkin, nd)[:3]
        ct_names, sn_names = ['ct0_2', 'ct2_2', 'ct4_2'], ['sn0', 'sn0_2', 'sn2_2', 'sn4_2']
        terms += eftlike_terms(kin, ells, nd, ct_names, sn_names)
        if solved:
            for n in ct_names: solved_list.append(dict(name=n, value=0., prior=dict(dist='norm', loc=0., scale=100.), marg=True))
            for n in sn_names: solved_list.append(dict(name=n, value=0., prior=dict(dist='norm', loc=0., scale=1000.), marg=True))
        else:
            for t in terms[3:]:
                t['factors'] = [(t.pop('solved'), 1)]
            params += ct_names + sn_names
    else:
        terms = kaiser_terms(ells, nkin, nd)
        params.append('sn0')
    W = window_matrix(kout, ells, kin, ells, width=0.006)
    shotnoise = 1. / nd
    shotnoisein = np.array([shotnoise * (ell == 0) for ell in ells])
    shotnoiseout = shotnoise * np.concatenate([np.full(nbins, float(ell == 0)) for ell in ells])
    offset = W.dot(np.repeat(shotnoisein, nkin)) - shotnoiseout
    obs = dict(name='LRG', theory=theory, terms=terms, nin=len(ells) * nkin, wmatrix=W, offset=offset, transform=transform,
               kout=kout, wmatrix_raw=W, shotnoise=shotnoise)
    if systematic:   # two systematic templates (reference window.py:470-473, :1253-1287): syst_0 sampled, syst_1 analytically marginalised
        t0 = np.concatenate([(ell == 0) * 3e3 * np.exp(-kout / 0.05) for ell in ells])
        t1 = np.concatenate([(ell == 2) * 1e3 * (kout / 0.1)**-1 for ell in ells])
        terms += [dict(scale=1., factors=[('syst_0', 1)], table=None, outvec=t0), dict(scale=1., factors=[], solved='syst_1', table=None, outvec=t1)]
        params.append('syst_0')
        solved_list.append(dict(name='syst_1', value=0., prior=dict(dist='norm', loc=0., scale=3.), marg=True))
        obs['systematic_templates'] = [t0, t1]
    spec = dict(params=params, observables=[obs], solved=solved_list)
    fidv = dict(_FID, ct0_2=0., ct2_2=0., ct4_2=0., sn0_2=0., sn2_2=0., sn4_2=0., syst_0=0.)
    spec = _finish(spec, [(kout, ells, nd)], rng, fidv)
    if systematic: spec['priors']['syst_0'] = dict(dist='uniform', limits=(-2., 2.))
    for n in params:
        if n.startswith('ct') or (n.startswith('sn') and n != 'sn0'):
            spec['priors'][n] = dict(dist='norm', loc=0., scale=10.)
    return spec
