"""CPU (NumPy/SciPy, FP64) restatement of the reference's per-point arithmetic on the batched full-shape path.

TEST INFRASTRUCTURE — the checker, never the product.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/.

Every function cites the reference lines it follows (paths relative to /root/reference).
Pinning status (tests/golden/make_golden.py):
  * window (P(k) and xi), chi2, prior, Kaiser (+ShapeFit, AP), EFT-like add-ons, damped BAO P(k) / xi wiring and broadband terms, the
    'cubic' observable transform, systematic templates, '.marg' / '.best' / '.prec' analytic marginalisation (the reference's own
    FastFisher / _solve with jax.jacfwd shimmed by an exact central difference), covariance -> precision with Hartlap / Percival factors,
    the velocileptors-style bias combination + interp1d, vmap conventions:
    PINNED to rounding against the reference's own NumPy path run under stub third-party modules
    (oracle/ref_harness.py) -> committed fixtures tests/golden/*.npz, and live in tests/test_oracle_vs_reference.py.
  * Taylor predict, MLP predict, FFTLog: the reference delegates these to packages absent here (cosmoprimo, jax)
    -> PARITY UNPINNED against those packages; pinned mathematically (Taylor of a polynomial is exact; the MLP operation vocabulary of
    emulators/conversion.py; FFTLog vs the defining integral and an analytic pair).
"""
import numpy as np
from scipy import special, interpolate


# --------------------------------------------------------------------------------------------------------------------
# mu quadrature: reference desilike/utils.py:625-643, desilike/theories/galaxy_clustering/base.py:201-208
# --------------------------------------------------------------------------------------------------------------------
def weights_mu(nmu):
    x, w = np.polynomial.legendre.leggauss(2 * nmu)
    return x[nmu:], (w[nmu:] + w[nmu - 1::-1]) / 2.


def legendre_weights(ells, nmu):
    mu, w = weights_mu(nmu)
    wmu = np.array([w * (2 * ell + 1) * special.legendre(ell)(mu) for ell in ells])
    return mu, wmu


# --------------------------------------------------------------------------------------------------------------------
# AP: reference desilike/theories/galaxy_clustering/base.py:211-223 (ap_k_mu), :325-353 (APEffect.calculate)
# --------------------------------------------------------------------------------------------------------------------
def ap_qpar_qper(mode, params, eta=1. / 3.):
    if mode == 'qiso':
        return params['qiso'], params['qiso']
    if mode == 'qap':
        qap = params['qap']
        return qap**(1 - eta), qap**(-eta)
    if mode == 'qisoqap':
        qiso, qap = params['qiso'], params['qap']
        return qiso * qap**(1 - eta), qiso * qap**(-eta)
    if mode == 'qparqper':
        return params['qpar'], params['qper']
    raise ValueError(mode)


def ap_k_mu(k, mu, qpar, qper):
    qap = qpar / qper
    jac = 1. / (qpar * qper**2)
    factorap = np.sqrt(1 + mu**2 * (1. / qap**2 - 1))
    kap = k[:, None] / qper * factorap
    muap = mu / qap / factorap
    return jac, kap, np.broadcast_to(muap, kap.shape)


# --------------------------------------------------------------------------------------------------------------------
# interpolation: reference desilike/jax.py:258-265.  The reference runs here only in its NumPy fallback
# (jax/interpax absent), i.e. scipy.interpolate.interp1d(kind='cubic', fill_value='extrapolate'): the not-a-knot
# C2 cubic spline.  This build adopts THAT mode (DESIGN.md "interpolation semantics").
# --------------------------------------------------------------------------------------------------------------------
def interp1d_cubic(xq, x, f):
    return interpolate.interp1d(x, f, kind='cubic', fill_value='extrapolate', axis=0)(xq)


def interp1d_cubic_local(xq, x, f):
    """The reference's jax branch (desilike/jax.py:258-261): ``interpax.interp1d(xq, x, f, method='cubic')`` -- interpax is a third-party
    dependency absent from /root/reference and from this image (unpinned in pyproject.toml:33-35), so this RESTATES its published algorithm
    (interpax/_spline.py: ``interp1d`` + ``approx_df(method='cubic')``), PARITY UNPINNED against interpax itself:
    C1 cubic Hermite interpolation on [x_i, x_i+1] with knot derivatives fx_i = (s_{i-1} + s_i) / 2 from the secant slopes s_i = (f_i+1 - f_i) /
    (x_i+1 - x_i) (fx_0 = s_0, fx_{n-1} = s_{n-2}), interval i = clip(searchsorted(x, xq, side='right'), 1, n - 1) - 1, and NaN outside
    [x_0, x_{n-1}] (interpax default ``extrap=False``)."""
    xq, x, f = np.asarray(xq, dtype='f8'), np.asarray(x, dtype='f8'), np.asarray(f, dtype='f8')
    s = np.diff(f, axis=0) / np.diff(x).reshape((-1,) + (1,) * (f.ndim - 1))
    fx = np.concatenate([s[:1], 0.5 * (s[:-1] + s[1:]), s[-1:]], axis=0)
    shape = xq.shape
    q = xq.ravel()
    i = np.clip(np.searchsorted(x, q, side='right'), 1, x.size - 1) - 1
    dx = x[i + 1] - x[i]
    t = (q - x[i]) / dx
    ex = (-1,) + (1,) * (f.ndim - 1)
    t, dx = t.reshape(ex), dx.reshape(ex)
    f0, f1, d0, d1 = f[i], f[i + 1], fx[i] * dx, fx[i + 1] * dx
    out = f0 + t * (d0 + t * ((3. * (f1 - f0) - 2. * d0 - d1) + t * (-2. * (f1 - f0) + d0 + d1)))
    out = np.where(((q < x[0]) | (q > x[-1])).reshape(ex), np.nan, out)
    return out.reshape(shape + f.shape[1:])


def interp1d_linear(xq, x, f):
    return interpolate.interp1d(x, f, kind='linear', fill_value='extrapolate', axis=0)(xq)


# --------------------------------------------------------------------------------------------------------------------
# templates: reference desilike/theories/galaxy_clustering/power_template.py:747-761 (ShapeFit), :372-376 (BAO)
# --------------------------------------------------------------------------------------------------------------------
def shapefit_factor(k_t, dm, dn, a=0.6, kp=0.03):
    return np.exp(dm / a * np.tanh(a * np.log(k_t / kp)) + dn * np.log(k_t / kp))


# --------------------------------------------------------------------------------------------------------------------
# Kaiser: reference desilike/theories/galaxy_clustering/full_shape.py:488-500 (tables), :545-550 (tracer combine),
# :628-634 (EFT-like counter / stochastic add-ons)
# --------------------------------------------------------------------------------------------------------------------
def kaiser_tables(k, mu, wmu, k_t, pk_t, f, qpar, qper, sigmapar=0., sigmaper=0., interp='scipy'):
    jac, kap, muap = ap_k_mu(k, mu, qpar, qper)
    sigmanl2 = kap**2 * (sigmapar**2 * muap**2 + sigmaper**2 * (1. - muap**2))
    damping = np.exp(-sigmanl2 / 2.)
    interp1d = interp1d_cubic if interp in ('scipy', 'not-a-knot') else interp1d_cubic_local   # reference jax.py:263-265 / :258-261
    pkmu = jac * damping * interp1d(np.log10(kap), np.log10(k_t), pk_t)

    def to_poles(x):
        return np.sum(x * wmu[:, None, :], axis=-1)

    return {'pk_dd': to_poles(pkmu), 'pk_dt': to_poles(f * muap**2 * pkmu), 'pk_tt': to_poles(f**2 * muap**4 * pkmu)}


def kaiser_tracer_power(tables, ells, b1X, b1Y, sn0, nd):
    sn = np.array([(ell == 0) for ell in ells], dtype='f8')[:, None] * sn0 / nd
    return b1X * b1Y * tables['pk_dd'] + (b1X + b1Y) * tables['pk_dt'] + tables['pk_tt'] + sn


def eftlike_matrices(k, ells, ct_names, sn_names, kp=1.):
    """``counterterm_matrix`` / ``stochastic_matrix`` [nell, nk, nparams] (reference full_shape.py:590-621);
    names are 'ct{ell}_{pow}', 'sn{ell}_{pow}' or 'sn0' (-> ell 0, constant)."""
    import re

    def matrix(names, base):
        mat = np.zeros((len(ells), len(k), len(names)), dtype='f8')
        for i, name in enumerate(names):
            if name == base + '0':
                ell, col = 0, np.ones_like(k)
            else:
                match = re.match(base + '(.*)_(.*)', name)
                ell, col = int(match.group(1)), (k / kp)**int(match.group(2))
            if ell in ells:
                mat[list(ells).index(ell), :, i] = col
        return mat

    return matrix(ct_names, 'ct'), matrix(sn_names, 'sn')


# --------------------------------------------------------------------------------------------------------------------
# damped BAO P(k, mu), model='standard': reference desilike/theories/galaxy_clustering/bao.py:117-134
# --------------------------------------------------------------------------------------------------------------------
def bao_damped_power(k, mu, wmu, k_t, pk_t, pknow_t, f_fid, qpar, qper, b1=1., dbeta=1., sigmas=0., sigmapar=9.,
                     sigmaper=6., mode='', smoothing_radius=15.):
    f = dbeta * f_fid
    jac, kap, muap = ap_k_mu(k, mu, qpar, qper)
    lk_t = np.log10(k_t)
    pknowap = interp1d_cubic(np.log10(kap), lk_t, pknow_t)
    pkap = interp1d_cubic(np.log10(kap), lk_t, pk_t)
    kk = k[:, None]
    pkwap = pkap - pknowap
    sigma_nl2ap = kap**2 * (sigmapar**2 * muap**2 + sigmaper**2 * (1. - muap**2))
    sk = 0.
    if mode == 'reciso': sk = np.exp(-1. / 2. * (kk * smoothing_radius)**2)
    Cap = (b1 + f * muap**2 * (1 - sk))**2 * np.exp(-sigma_nl2ap / 2.)
    fog = 1. / (1. + (sigmas * kk * mu)**2 / 2.)**2.
    B = (b1 + f * mu**2 * (1 - sk))**2 * fog
    pknow = interp1d_cubic(np.log10(kk), lk_t, pknow_t)
    pkmu = B * pknow + Cap * pkwap
    return np.sum(pkmu * wmu[:, None, :], axis=-1)


# --------------------------------------------------------------------------------------------------------------------
# P -> xi: reference desilike/theories/galaxy_clustering/base.py:62-77 (constants), :127-136 (get_corr)
# --------------------------------------------------------------------------------------------------------------------
class PowerToCorr(object):

    def __init__(self, s, ells, kin=None, interp_order=1):
        from .fftlog import PowerToCorrelation
        self.s = np.asarray(s, dtype='f8')
        self.ells = tuple(ells)
        self.k = np.logspace(-4., 3., 2048)
        self.kin = np.geomspace(self.k[0], 0.6, int(300. / interp_order + 0.5)) if kin is None else np.asarray(kin, dtype='f8')
        mask = self.k > self.kin[-1]
        self.logk_high = np.log10(self.k[mask] / self.kin[-1])
        self.damp_high = np.exp(-(self.k[mask] / self.kin[-1] - 1.)**2 / (2. * (10.)**2))
        self.k_mid = self.k[~mask]
        self.interp_order = interp_order
        self.fftlog = PowerToCorrelation(self.k, ell=self.ells, q=0, lowring=True)

    def extend(self, power):
        """(nell, nkin) -> (nell, 2048): interpolation onto the FFTLog grid + high-k extrapolation (:127-133)."""
        tmp = []
        interp = interp1d_linear if self.interp_order == 1 else interp1d_cubic
        for pk in power:
            slope_high = (pk[-1] - pk[-2]) / np.log10(self.kin[-1] / self.kin[-2])
            mid = interp(np.log10(self.k_mid), np.log10(self.kin), pk)
            tmp.append(np.concatenate([mid, (pk[-1] + slope_high * self.logk_high) * self.damp_high], axis=-1))
        return np.vstack(tmp)

    def __call__(self, power):
        s, corr = self.fftlog(self.extend(power))
        return np.array([np.interp(self.s, ss, cc) for ss, cc in zip(s, corr)])


# --------------------------------------------------------------------------------------------------------------------
# emulators.  Taylor layout: reference desilike/emulators/__init__.py:471-507 (center, powers[T, d], derivatives[T, *yshape]
# with 1/alpha! folded in); predict is cosmoprimo's (absent) -> restated:  y = tensordot(derivatives, prod(diffs**powers)).
# MLP: operation vocabulary reference desilike/emulators/conversion.py:18-33,76-104.
# --------------------------------------------------------------------------------------------------------------------
def taylor_terms(ndim, order):
    """Enumeration order of reference emulators/__init__.py:485-503 for a uniform max order per parameter."""
    import itertools
    powers = []
    for o in range(order + 1):
        for indices in itertools.product(range(ndim), repeat=o):
            orders = tuple(np.bincount(indices, minlength=ndim).astype('i4')) if o else (0,) * ndim
            if orders not in powers:
                powers.append(orders)
    return np.array(powers, dtype='i4')


def taylor_monomials(X, center, powers):
    """X [N, d] -> [N, T]."""
    diffs = np.asarray(X, dtype='f8') - center
    return np.prod(np.power(diffs[:, None, :], powers[None, :, :]), axis=-1)


def taylor_predict(X, center, powers, derivatives):
    """X [N, d] -> [N, *yshape]."""
    mono = taylor_monomials(X, center, powers)
    return np.tensordot(mono, derivatives, axes=(1, 0))


def taylor_fit_polynomial(func, center, ndim, order, delta):
    """Taylor coefficients d^alpha y / alpha! at ``center`` by tensor-product central finite differences of stencil
    width ``order + 1 (+1)`` -- exact for polynomials of degree <= order; used to build config C2(i) coefficients."""
    from numpy.polynomial import polynomial as npoly
    powers = taylor_terms(ndim, order)
    npts = order + 1 + (order % 2 == 0) * 0
    if npts % 2 == 0: npts += 1
    nodes = (np.arange(npts) - npts // 2).astype('f8')
    # 1-D matrix mapping samples at nodes -> polynomial coefficients c_j (y = sum_j c_j t^j)
    V = np.vander(nodes, npts, increasing=True)
    Vinv = np.linalg.inv(V)
    grids = np.meshgrid(*[nodes] * ndim, indexing='ij')
    pts = np.stack([g.ravel() for g in grids], axis=-1) * delta + center
    Y = func(pts)  # [npts**ndim, ny]
    Y = Y.reshape((npts,) * ndim + Y.shape[1:])
    for ax in range(ndim):
        Y = np.moveaxis(np.tensordot(Vinv, np.moveaxis(Y, ax, 0), axes=(1, 0)), 0, ax)
    derivs = []
    for p in powers:
        derivs.append(Y[tuple(p)] / np.prod(np.asarray(delta, dtype='f8')**p))
    return powers, np.array(derivs)


_ACTIVATIONS = {
    'silu': lambda v: v / (1 + np.exp(-v)),
    'relu': lambda v: np.maximum(v, 0.),
    'tanh': np.tanh,
    'identity': lambda v: v,
}


def mlp_predict(X, mlp):
    """X [N, d] -> [N, ny].  ``mlp`` = dict(xlimits [d, 2] or None, layers [(kernel, bias), ...], activation,
    ylimits [ny, 2] or None (inverse min-max scaling), pca (eigvecs [npc, ny], mean [ny]) or None)."""
    v = np.asarray(X, dtype='f8')
    if mlp.get('xlimits') is not None:
        lim = mlp['xlimits']
        v = (v - lim[..., 0]) / (lim[..., 1] - lim[..., 0])
    act = _ACTIVATIONS[mlp.get('activation', 'silu')]
    nl = len(mlp['layers'])
    for il, (kernel, bias) in enumerate(mlp['layers']):
        v = (v[..., None, :] @ kernel)[..., 0, :] + bias
        if il < nl - 1:
            v = act(v)
    if mlp.get('pca') is not None:
        eigvecs, mean = mlp['pca']
        v = v @ eigvecs + mean
    if mlp.get('ylimits') is not None:
        lim = mlp['ylimits']
        v = v * (lim[..., 1] - lim[..., 0]) + lim[..., 0]
    return v


# --------------------------------------------------------------------------------------------------------------------
# window: reference desilike/observables/galaxy_clustering/window.py:445-473 (P), :717-728 (xi)
# --------------------------------------------------------------------------------------------------------------------
def window_apply(power, wmatrix=None, offset=None, kmask=None, shotnoisein=None, shotnoiseout=None):
    """power [..., nell_in, nk_in] -> flat [..., nobs]."""
    theory = power if shotnoisein is None else power + shotnoisein[:, None]
    theory = theory.reshape(theory.shape[:-2] + (-1,))
    if wmatrix is not None:
        theory = theory @ wmatrix.T
    if offset is not None:
        theory = theory + offset
    if kmask is not None:
        theory = theory[..., kmask]
    if shotnoiseout is not None:
        theory = theory - shotnoiseout
    return theory


def cubic_transform(flattheory, flatdata):
    """reference desilike/observables/galaxy_clustering/power_spectrum.py:402-404."""
    return (3. * np.cbrt(flattheory / flatdata) - 2.) * flatdata


# --------------------------------------------------------------------------------------------------------------------
# velocileptors-style table models: reference desilike/theories/galaxy_clustering/full_shape.py:1182-1186 (bias combine),
# :1295-1313 (LPTVelocileptorsTracerPowerSpectrumMultipoles.calculate, standard bias basis: cubic re-interpolation pt.k -> k)
# --------------------------------------------------------------------------------------------------------------------
LPT_BIAS_PARAMS = ['b1', 'b2', 'bs', 'b3', 'alpha0', 'alpha2', 'alpha4', 'alpha6', 'sn0', 'sn2', 'sn4']


def lpt_physical_pars(p, sigma8, fsigma8, shotnoise=1e4, fsat=0.1, sigv=5., nd=1e-4):
    """Physical prior basis of the velocileptors tracers, restated from reference full_shape.py:1299-1307 (+ :1154-1157: snd = shotnoise * nd with
    nd = 1e-4): sampled (b1p, b2p, bsp, b3p, alpha0p, alpha2p, alpha4p, sn0p, sn2p, sn4p) -> the 11 standard-basis parameters fed to
    ``combine_bias_terms_poles``."""
    f = fsigma8 / sigma8
    snd = shotnoise * nd
    b1L, b2L, bsL, b3L = p['b1p'] / sigma8 - 1., p['b2p'] / sigma8**2, p['bsp'] / sigma8**2, p['b3p'] / sigma8**3
    pars = [b1L, b2L, bsL, b3L]
    pars += [(1 + b1L)**2 * p['alpha0p'], f * (1 + b1L) * (p['alpha0p'] + p['alpha2p']), f * (f * p['alpha2p'] + (1 + b1L) * p['alpha4p']), f**2 * p['alpha4p']]
    pars += [p['sn{:d}p'.format(i)] * snd * (fsat if i > 0 else 1.) * sigv**i for i in [0, 2, 4]]
    return pars


def lpt_bias_monomials(pars, nd=1e-4):
    b1, b2, bs, b3, alpha0, alpha2, alpha4, alpha6, sn0, sn2, sn4 = pars
    return np.array([1, b1, b1**2, b2, b1 * b2, b2**2, bs, b1 * bs, b2 * bs, bs**2, b3, b1 * b3, alpha0, alpha2, alpha4, alpha6,
                     sn0 / nd, sn2 / nd, sn4 / nd])


def tablevel_combine_bias_terms_poles(pktable, pars, nd=1e-4):
    """pktable [nell, nk, 19] -> P_ell(k) [nell, nk]."""
    return np.sum(pktable * lpt_bias_monomials(pars, nd=nd), axis=-1)


def lpt_tracer_power(pktable, kpt, k, pars, nd=1e-4):
    """reference full_shape.py:1310-1312: ``interp1d(self.k, self.pt.k, combine(...)[index].T).T`` (cubic; SciPy not-a-knot in the NumPy fallback)."""
    return interp1d_cubic(k, kpt, tablevel_combine_bias_terms_poles(pktable, pars, nd=nd).T).T


# --------------------------------------------------------------------------------------------------------------------
# chi2: reference desilike/likelihoods/base.py:13-17, :658-664
# --------------------------------------------------------------------------------------------------------------------
def chi2(flatdiff, precision):
    if precision.ndim == 1:
        return np.sum(flatdiff * precision * flatdiff, axis=-1)
    return np.einsum('...i,ij,...j->...', flatdiff, precision, flatdiff)


# --------------------------------------------------------------------------------------------------------------------
# prior: reference desilike/parameter.py:1889-1897, :1994-2017 (uniform / normal fast paths, zero-lag removed)
# --------------------------------------------------------------------------------------------------------------------
def logprior_one(x, prior):
    x = np.asarray(x, dtype='f8')
    if prior is None:
        return np.zeros_like(x)
    lo, hi = prior.get('limits', (-np.inf, np.inf))
    isin = (lo <= x) & (x <= hi)
    if prior.get('dist', 'uniform') == 'norm':
        with np.errstate(over='ignore'):
            val = -0.5 * (x - prior['loc'])**2 / prior['scale']**2
        return np.where(isin, val, -np.inf)
    return np.where(isin, 0., -np.inf)


def logprior(X, names, priors):
    """X [N, d]; priors: {name: prior dict or None}."""
    X = np.asarray(X, dtype='f8')
    out = np.zeros(X.shape[0], dtype='f8')
    for i, name in enumerate(names):
        out = out + logprior_one(X[:, i], priors.get(name))
    return out


# --------------------------------------------------------------------------------------------------------------------
# analytic marginalisation: reference desilike/likelihoods/base.py:129-200 (FastFisher.__call__), :314-413 (_solve).
# jax.jacfwd is replaced by the explicit Jacobian (the theory is exactly linear in the solved parameters).
# --------------------------------------------------------------------------------------------------------------------
def solve_marginalised(flatdiff0, flatderiv, precisions, x0, solved_priors, marg_mask, with_hessian=False):
    """
    One point.  ``flatdiff0`` list (per likelihood) of d0 = t(theta, x0) - data, ``flatderiv`` list of J [ns, n_L] (rows of zeros
    for solved params a likelihood does not depend on), ``precisions`` list of [n_L, n_L] or [n_L];
    returns dict(loglikelihood, logprior_solved, x, dx, hessian) following _solve term by term.
    """
    ns = len(x0)
    x0 = np.asarray(x0, dtype='f8')
    like_grad, like_hess = [], []
    for d0, J, P in zip(flatdiff0, flatderiv, precisions):
        derivp = J * P if P.ndim == 1 else J.dot(P)
        like_grad.append(-derivp.dot(d0))
        like_hess.append(-derivp.dot(J.T))
    prior_grad, prior_hess = np.zeros(ns), np.zeros(ns)
    for i, prior in enumerate(solved_priors):
        loc = prior.get('loc', 0.) if prior else 0.
        scale = prior.get('scale', np.inf) if prior else np.inf
        prec = scale**(-2)
        prior_grad[i] = -(x0[i] - loc) * prec
        prior_hess[i] = -prec
    grad = prior_grad + sum(like_grad)
    hess = np.diag(prior_hess) + sum(like_hess)
    dx = -np.linalg.solve(hess, grad)
    x = x0 + dx
    logprior_solved = 0.
    for i, prior in enumerate(solved_priors):
        logprior_solved += float(logprior_one(x[i], prior))
    loglike = 0.
    for d0, P, g, H in zip(flatdiff0, precisions, like_grad, like_hess):
        loglike += -0.5 * chi2(d0, P)
        loglike += 0.5 * dx.dot(H).dot(dx) + g.dot(dx)
    marg = np.flatnonzero(marg_mask)
    if marg.size:
        loglike += -0.5 * np.linalg.slogdet(-hess[np.ix_(marg, marg)])[1]
    out = dict(loglikelihood=loglike, logprior_solved=logprior_solved, x=x, dx=dx)
    if with_hessian:
        out['hessian'] = hess
    return out
