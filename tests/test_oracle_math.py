"""Mathematical pins of the oracle restatements whose reference implementation lives in absent third-party packages
(cosmoprimo Taylor/MLP predict, FFTLog; jax): SURVEY.md section 8c."""
import numpy as np

from oracle import core, configs, pipeline
from oracle.fftlog import PowerToCorrelation, brute_force_xi


def test_fftlog_analytic_gaussian_pair():
    k = np.logspace(-4, 3, 2048)
    a = 5.
    fft = PowerToCorrelation(k, ell=(0, 2, 4), q=0, lowring=True)
    pk = np.exp(-0.5 * (k * a)**2)
    s, xi = fft(np.array([pk, pk, pk]))
    true = (2 * np.pi * a**2)**-1.5 * np.exp(-s[0]**2 / (2 * a**2))
    m = (s[0] > 1) & (s[0] < 20)
    assert np.max(np.abs(xi[0][m] / true[m] - 1)) < 1e-6


def test_fftlog_vs_defining_integral():
    """reference theories/galaxy_clustering/base.py:163-168."""
    from oracle import synthetic
    k = np.logspace(-4, 3, 2048)
    pk = synthetic.pk_fid(k) * np.exp(-(k / 2.)**2)
    fft = PowerToCorrelation(k, ell=(0, 2, 4), q=0, lowring=True)
    s, xi = fft(np.array([pk, pk, pk]))
    for i, ell in enumerate((0, 2, 4)):
        idx = np.searchsorted(s[i], [25., 60., 110.])
        bf = brute_force_xi(k, pk, ell, s[i][idx])
        assert np.allclose(xi[i][idx], bf, rtol=2e-4, atol=1e-8)


def test_fftlog_grid_and_padding():
    k = np.logspace(-4, 3, 2048)
    fft = PowerToCorrelation(k, ell=(0, 2), q=0, lowring=True)
    assert fft.padded_size == 4096 and fft.pad_in_left == 1024
    assert np.all(np.diff(fft.s, axis=-1) > 0)
    assert np.allclose(fft.s[0] * k[::-1], fft.s[0][0] * k[-1])


def test_taylor_exact_for_polynomial():
    rng = np.random.default_rng(0)
    ndim, order = 3, 3
    powers = core.taylor_terms(ndim, order)
    assert len(powers) == 20  # C(3+3, 3)
    coeffs = rng.standard_normal((len(powers), 5))
    center = np.array([0.3, -0.2, 1.1])

    def poly(X):
        return core.taylor_predict(X, center, powers, coeffs)

    p2, d2 = core.taylor_fit_polynomial(poly, center, ndim, order, np.array([0.1, 0.2, 0.05]))
    assert np.array_equal(p2, powers)
    assert np.allclose(d2, coeffs, rtol=1e-8, atol=1e-9)
    X = center + 0.3 * rng.standard_normal((7, ndim))
    assert np.allclose(core.taylor_predict(X, center, p2, d2), poly(X), rtol=1e-8)


def test_taylor_term_count_order4():
    assert len(core.taylor_terms(4, 4)) == 70 and len(core.taylor_terms(6, 4)) == 210


def test_mlp_predict_matches_manual():
    rng = np.random.default_rng(1)
    mlp = configs.mlp_engine(rng, 12, hidden=(8, 8), kin=np.linspace(0.01, 0.2, 4))
    X = rng.uniform(0.95, 1.05, (5, 4)); X[:, 2] -= 1.
    v = (X - mlp['xlimits'][:, 0]) / (mlp['xlimits'][:, 1] - mlp['xlimits'][:, 0])
    for i, (W, b) in enumerate(mlp['layers']):
        v = v @ W + b
        if i < 2: v = v / (1 + np.exp(-v))
    v = v * (mlp['ylimits'][:, 1] - mlp['ylimits'][:, 0]) + mlp['ylimits'][:, 0]
    assert np.allclose(core.mlp_predict(X, mlp), v, rtol=1e-14)


def test_marginalisation_equals_profile_and_gaussian_integral():
    """'.best' logL equals the numerically maximised full logL over x; '.marg' = '.best' - 1/2 log det(-H)
    (reference likelihoods/base.py:385-404)."""
    from scipy import optimize
    spec = configs.config_kaiser_window(eftlike=True, solved=True, seed=3)
    X = configs.sample_points(spec, 3, seed=9)
    marg = pipeline.evaluate(spec, X)
    for s in spec['solved']: s['marg'] = False
    best = pipeline.evaluate(spec, X)
    full = configs.config_kaiser_window(eftlike=True, solved=False, seed=3)
    names = [s['name'] for s in spec['solved']]
    for n in names: full['priors'][n] = [s['prior'] for s in spec['solved'] if s['name'] == n][0]
    order = [full['params'].index(n) for n in names]
    for ip in range(3):
        def neg(x):
            p = np.zeros(len(full['params']))
            p[:5] = X[ip]
            p[order] = x
            return -pipeline.evaluate(full, p[None])['logposterior'][0]
        res = optimize.minimize(neg, best['x_solved'][ip], method='BFGS', options=dict(gtol=1e-10))
        assert abs(-res.fun - best['logposterior'][ip]) < 1e-6
        assert abs(neg(best['x_solved'][ip]) + best['logposterior'][ip]) < 1e-8
    P, obs = spec['precision'][0], spec['observables'][0]
    from oracle.pipeline import observable_flat
    Xn = X
    J = np.stack([observable_flat(obs, Xn, spec['params'], {}, only_solved=n) for n in names], axis=1)
    prec = np.array([1. / s['prior']['scale']**2 for s in spec['solved']])
    for ip in range(3):
        H = J[ip] @ P @ J[ip].T + np.diag(prec)
        assert abs((best['loglikelihood'][ip] - 0.5 * np.linalg.slogdet(H)[1]) - marg['loglikelihood'][ip]) < 1e-8


def test_legendre_weights_sum_to_one():
    """reference theories/tests/test_galaxy_clustering.py:6-12."""
    mu, w = core.weights_mu(8)
    assert abs(w.sum() - 1.) < 1e-14 and np.all((mu > 0) & (mu < 1))


def test_kaiser_closed_form():
    """Kaiser multipoles (b^2 + 2/3 b f + f^2/5, 4/3 b f + 4/7 f^2, 8/35 f^2) P(k) at q = 1."""
    spec = configs.config_c1()
    obs = spec['observables'][0]
    th = obs['theory']
    X = np.array([[1., 1., 0., 1., 2., 0.]])
    flat = pipeline.evaluate(spec, X, return_theory=True)['flattheory'][0].reshape(3, -1)
    from oracle import synthetic
    pk, f, b = synthetic.pk_fid(th['k']), synthetic.F_FID, 2.
    expected = np.array([(b**2 + 2. / 3. * b * f + f**2 / 5.) * pk, (4. / 3. * b * f + 4. / 7. * f**2) * pk, 8. / 35. * f**2 * pk])
    assert np.max(np.abs(flat / expected - 1.)) < 5e-7  # spline interpolation error of the wiggly synthetic P(k)
    assert abs(pipeline.evaluate(spec, X)['loglikelihood'][0]) < 1e-18


def test_local_cubic_interpolant_properties():
    """oracle/core.py:interp1d_cubic_local restates interpax's C1 'cubic' (the reference's jax branch, jax.py:258-261; interpax is absent here):
    mathematical pins -- interpolates the knots, reproduces quadratics exactly on a uniform grid (mean of secant slopes = exact derivative),
    equals SciPy's CubicHermiteSpline fed with the same knot slopes, is C1 across knots, NaN outside the knot range."""
    from scipy.interpolate import CubicHermiteSpline
    from oracle import core
    rng = np.random.default_rng(0)
    x = np.linspace(-1., 2., 40)
    f = np.stack([np.sin(3. * x), 0.3 * x**2 - x + 2.], axis=-1)
    assert np.allclose(core.interp1d_cubic_local(x, x, f), f, rtol=0, atol=1e-14)
    xq = rng.uniform(x[1], x[-2], 200)   # (the end intervals use one-sided slopes)
    got = core.interp1d_cubic_local(xq, x, f)
    assert np.allclose(got[:, 1], 0.3 * xq**2 - xq + 2., rtol=1e-12)
    xn = np.sort(rng.uniform(0., 1., 30)); fn = np.cos(5. * xn)
    s = np.diff(fn) / np.diff(xn)
    fx = np.concatenate([s[:1], 0.5 * (s[:-1] + s[1:]), s[-1:]])
    xq = rng.uniform(xn[0], xn[-1], 300)
    assert np.allclose(core.interp1d_cubic_local(xq, xn, fn), CubicHermiteSpline(xn, fn, fx)(xq), rtol=1e-12, atol=1e-14)
    eps = 1e-7
    left, right = core.interp1d_cubic_local(xn[5:-5] - eps, xn, fn), core.interp1d_cubic_local(xn[5:-5] + eps, xn, fn)
    assert np.allclose((right - left) / (2. * eps), fx[5:-5], rtol=1e-4, atol=1e-4)
    assert np.all(np.isnan(core.interp1d_cubic_local(np.array([xn[0] - 0.1, xn[-1] + 0.1]), xn, fn)))
