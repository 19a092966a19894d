"""GPU parity of the theory generators (Kaiser + ShapeFit + AP, MLP) and of the CUDA path against the committed golden
fixtures, i.e. against outputs of the reference's OWN code (tests/golden/make_golden.py).  Run with ``-m gpu``."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import configs, pipeline

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(f for f in glob.glob(os.path.join(HERE, 'golden', '*.npz')) if 'bao_theory' not in f and 'bao_broadband' not in f and 'lpt_combine' not in f and 'xi_window' not in f and 'covariance_corrections' not in f and 'fiber_collisions' not in f and 'taylor_state' not in f and 'kaiser_corr' not in f)


def test_kaiser_tables_match_oracle(gpu_required):
    """Feature-level parity: pk_dd / pk_dt / pk_tt tables, relative 1e-10 (FP64)."""
    import desilike_b200
    spec = configs.config_kaiser_window(eftlike=True, solved=True)
    plan = desilike_b200.Plan(spec, merge_terms=False)   # the four plain tables as separate output segments
    X = configs.sample_points(spec, 77, seed=3)
    feats = plan.features(0, X)
    gen = plan.lowered.gens[0]
    tabs = pipeline.theory_tables(spec['observables'][0]['theory'], X, spec['params'])
    assert np.all(feats[:, 0] == 1.) and np.all(feats[:, 1] == 0.)
    for name in ['pk_dd', 'pk_dt', 'pk_tt', 'pk11_l0']:
        foff, K, _ = gen.tables[name]
        got, ref = feats[:, foff:foff + K], tabs[name]
        assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=0)) < 1e-10, name
    plan.close()


@pytest.mark.parametrize('fn', FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_cuda_matches_reference_fixture(gpu_required, fn):
    """CUDA path vs the reference's own outputs (Kaiser + ShapeFit, window, EFT-like terms, .marg / .best)."""
    import desilike_b200
    data = np.load(fn)
    case = json.loads(str(data['case']))
    spec = getattr(configs, case['config'])(**case['kwargs'])
    if case['mode'] == '.best':
        for s in spec['solved']: s['marg'] = False
    if case['mode'] == '.prec':   # pre-marginalised precision (reference likelihoods/base.py:257-312), Jacobian evaluated on the GPU at plan time
        for s in spec['solved']: s['mode'] = 'prec'
        spec['fiducial'] = case['fiducial']
    plan = desilike_b200.Plan(spec)
    X = data['X']
    got = plan.eval(X)
    theory = plan.theory(X)
    scale = np.max(np.abs(data['flattheory']), axis=0)
    assert np.max(np.abs(theory - data['flattheory']) / scale) < 1e-10
    for key in ['logposterior', 'loglikelihood', 'logprior']:
        ref = data[key]
        finite = np.isfinite(ref)
        if key == 'logposterior':
            assert np.array_equal(np.isneginf(ref), np.isneginf(got[key]))
        if key == 'logprior':
            finite &= np.isfinite(got[key])
        atol = 2e-6 if 'cubic_marg' in fn else 1e-7   # (finite-difference Jacobian in that fixture, see tests/test_oracle_golden.py)
        assert np.allclose(got[key][finite], ref[finite], rtol=1e-9, atol=atol), (key, np.max(np.abs(got[key][finite] - ref[finite])))
    plan.close()


def test_c1_full_batch(gpu_required):
    """BASELINE configs[0]: Kaiser + ShapeFit -> chi2, 1024 points."""
    import desilike_b200
    spec = configs.config_c1()
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 1024, seed=8)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-6)
    assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood']) / np.maximum(1., np.abs(ref['loglikelihood']))) < 1e-9
    plan.close()


def test_kaiser_with_damping_and_apmodes(gpu_required):
    import desilike_b200
    for apmode, pnames in [('qisoqap', ['qiso', 'qap']), ('qiso', ['qiso']), ('qap', ['qap'])]:
        spec = configs.config_c1()
        th = spec['observables'][0]['theory']
        th['template'] = dict(th['template'], apmode=apmode)
        th['params'] = pnames + ['dm', 'df', 'sigmapar', 'sigmaper']
        th['defaults'] = dict(th['defaults'], qiso=1., qap=1.)
        spec['params'] = pnames + ['dm', 'df', 'sigmapar', 'sigmaper', 'b1', 'sn0']
        spec['priors'] = {}
        rng = np.random.default_rng(4)
        X = np.column_stack([rng.uniform(0.95, 1.05, 50) for _ in pnames] + [rng.uniform(-0.05, 0.05, 50), rng.uniform(0.9, 1.1, 50),
                            rng.uniform(0., 5., 50), rng.uniform(0., 5., 50), rng.uniform(1.5, 2.5, 50), rng.normal(0., 0.1, 50)])
        plan = desilike_b200.Plan(spec)
        ref = pipeline.evaluate(spec, X, return_theory=True)
        assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10, apmode
        assert np.allclose(plan.eval(X)['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-6)
        plan.close()


def test_mlp_emulated_marginalised(gpu_required):
    """C4-like: MLP pt emulator -> EFT-like combine -> window -> chi2 with 7 marginalised terms per tracer."""
    import desilike_b200
    spec = configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(16, 16))
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 150, seed=14)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    assert np.max(np.abs(got['x_solved'] - ref['x_solved']) / np.maximum(1., np.abs(ref['x_solved']))) < 1e-7
    plan.close()


def test_mlp_full_width(gpu_required):
    """Reference-like MLP width (3 x 64, SiLU) at BASELINE C4 table sizes for one tracer; tanh / relu variants."""
    import desilike_b200
    for act in ['silu', 'tanh', 'relu']:
        spec = configs.config_c4(ntracers=1, nkin=350, nbins=36, hidden=(64, 64, 64))
        spec['observables'][0]['theory']['mlp']['activation'] = act
        from oracle.pipeline import observable_flat
        plan = desilike_b200.Plan(spec)
        X = configs.sample_points(spec, 200, seed=15)
        ref = pipeline.evaluate(spec, X, return_theory=True)
        got = plan.eval(X)
        assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10, act
        assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8, act
        plan.close()


@pytest.mark.parametrize('hidden', [(128, 96), (512, 512, 512, 512), (64, 64, 64, 64, 64)])
def test_mlp_wide_networks(gpu_required, hidden):
    """MLP emulators beyond the fused kernel's limits (SURVEY.md section 8d stress variant 4 x 512; widths > 64 run layer by layer as output-mode
    launches of the contraction kernel, resident or streaming A) -- and a deeper 64-wide one through the fused kernel."""
    import desilike_b200
    spec = configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=hidden)
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 300, seed=16)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    assert np.max(np.abs(got['x_solved'] - ref['x_solved']) / np.maximum(1., np.abs(ref['x_solved']))) < 1e-7
    plan.close()


@pytest.mark.parametrize('fftlog', ['fold', 'kernel'])
def test_c3_bao_fftlog(gpu_required, fftlog):
    """BASELINE configs[2]: damped-BAO P(k) -> FFTLog -> xi_ell(s) (+ marginalised broadband) -> chi2; with the P -> xi operator folded
    into the term matrices (default) and through the explicit FFTLog kernel."""
    import desilike_b200
    spec = configs.config_c3()
    spec['observables'][0]['theory']['fftlog'] = fftlog
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 96, seed=33)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    theory = plan.theory(X)
    assert np.max(np.abs(theory - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood'])) < 5e-5
    assert np.max(np.abs(got['x_solved'] - ref['x_solved'])) < 1e-9
    plan.close()


def test_c3_fixed_broadband_and_reciso(gpu_required):
    import desilike_b200
    spec = configs.config_c3(broadband_solved=False)
    spec['observables'][0]['theory']['mode'] = 'reciso'
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 40, seed=34)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.allclose(plan.eval(X)['logposterior'], ref['logposterior'], rtol=1e-8, atol=1e-6)
    plan.close()


def test_bao_power_spectrum(gpu_required):
    """Damped BAO P_ell(k) observable (reference bao.py:117-151) with a power-law broadband (bao.py:527-529)."""
    import desilike_b200
    rng = np.random.default_rng(5)
    k = np.arange(0.02, 0.3, 0.005)
    ells = (0, 2)
    tpl = configs.bao_template()
    theory = dict(kind='bao_pk', k=k, ells=ells, nmu=10, template=tpl, mode='', params=['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper'],
                  defaults=dict(qpar=1., qper=1., df=1.))
    terms = [dict(scale=1., factors=[], table='power')]
    solved = []
    kp = 2. * np.pi / 100.
    for ill, ell in enumerate(ells):
        for power in (-1, 0, 1):
            cv = np.zeros((len(ells), k.size)); cv[ill] = (k / kp)**power
            name = 'al{:d}_{:d}'.format(ell, power)
            terms.append(dict(scale=1., factors=[], solved=name, table=None, colvec=cv.ravel()))
            solved.append(dict(name=name, value=0., prior=dict(dist='norm', loc=0., scale=1e4), marg=True))
    obs = dict(name='pk', theory=theory, terms=terms, nin=len(ells) * k.size, wmatrix=None, offset=None)
    spec = dict(params=list(theory['params']), observables=[obs], solved=solved)
    fid = np.array([[1., 1., 2., 1., 4., 6., 4.]])
    obs['data'] = pipeline.observable_flat(obs, fid, spec['params'], {s['name']: 0. for s in solved})[0]
    spec['precision'] = [1. / (0.05 * np.abs(obs['data']) + 20.)**2]   # diagonal precision
    spec['priors'] = configs.default_priors(spec['params'])
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 64, seed=35)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    plan.close()


@pytest.mark.parametrize('fftlog', ['fold', 'kernel'])
def test_bao_kernels_match_reference_classes(gpu_required, fftlog):
    """GPU damped-BAO P_ell(k) and FFTLog xi_ell(s) vs the outputs of the reference's own classes (tests/golden/bao_theory.npz)."""
    import desilike_b200
    from oracle import synthetic
    data = np.load(os.path.join(HERE, 'golden', 'bao_theory.npz'))
    X, k, s, ells = data['X'], data['k'], data['s'], (0, 2)
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper']
    for which, kt in [('power', data['k_template_pk']), ('corr', data['k_template_xi'])]:
        tpl = dict(kind='bao', k=kt, pk_dd_fid=synthetic.pk_fid(kt), pknow_dd_fid=synthetic.pknow_fid(kt), f_fid=synthetic.F_FID, apmode='qparqper', eta=1. / 3.)
        theory = dict(ells=ells, nmu=10, template=tpl, mode='', params=names, defaults={}, fftlog=fftlog)
        theory.update(dict(kind='bao_pk', k=k) if which == 'power' else dict(kind='bao_xi', s=s))
        n = 2 * (k.size if which == 'power' else s.size)
        obs = dict(name='x', theory=theory, terms=[dict(scale=1., factors=[], table=which)], nin=n, wmatrix=None, offset=None, data=np.zeros(n))
        spec = dict(params=names, priors={}, solved=[], observables=[obs], precision=[np.ones(n)])
        plan = desilike_b200.Plan(spec)
        got = plan.theory(X).reshape(data[which].shape)
        assert np.max(np.abs(got - data[which])) <= 1e-10 * np.max(np.abs(data[which])), which
        plan.close()


def test_nonuniform_template_grids(gpu_required):
    """Template grids that are NOT uniform in log10 k take the searched-interval spline path of the Kaiser / BAO kernels (the closed-form
    interval lookup is only enabled when spline.is_uniform holds at plan time): parity with the oracle on a jittered grid."""
    import desilike_b200
    from oracle import synthetic
    from desilike_b200.spline import is_uniform
    rng = np.random.default_rng(17)
    # Kaiser + window
    spec = configs.config_kaiser_window(eftlike=False, solved=False)
    tpl = spec['observables'][0]['theory']['template']
    k = np.sort(tpl['k'] * np.exp(0.3 * np.log(tpl['k'][1] / tpl['k'][0]) * rng.uniform(-1., 1., tpl['k'].size)))
    assert is_uniform(np.log10(tpl['k'])) and not is_uniform(np.log10(k))
    tpl.update(k=k, pk_dd_fid=synthetic.pk_fid(k), pknow_dd_fid=synthetic.pknow_fid(k))
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 64, seed=18)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    plan.close()
    # damped BAO P(k)
    kk = np.arange(0.02, 0.3, 0.005)
    tplb = configs.bao_template()
    kb = np.sort(tplb['k'] * np.exp(0.3 * np.log(tplb['k'][1] / tplb['k'][0]) * rng.uniform(-1., 1., tplb['k'].size)))
    tplb.update(k=kb, pk_dd_fid=synthetic.pk_fid(kb), pknow_dd_fid=synthetic.pknow_fid(kb))
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper']
    theory = dict(kind='bao_pk', k=kk, ells=(0, 2), nmu=10, template=tplb, mode='', params=names, defaults={})
    n = 2 * kk.size
    obs = dict(name='x', theory=theory, terms=[dict(scale=1., factors=[], table='power')], nin=n, wmatrix=None, offset=None, data=np.zeros(n))
    specb = dict(params=names, priors={}, solved=[], observables=[obs], precision=[np.ones(n)])
    Xb = np.column_stack([rng.uniform(0.95, 1.05, 32), rng.uniform(0.95, 1.05, 32), rng.uniform(1.5, 2.5, 32), rng.uniform(0.8, 1.2, 32), rng.uniform(2., 8., 32),
                          rng.uniform(2., 8., 32), rng.uniform(2., 8., 32)])
    planb = desilike_b200.Plan(specb)
    refb = pipeline.evaluate(specb, Xb, return_theory=True)
    assert np.max(np.abs(planb.theory(Xb) - refb['flattheory'])) < 1e-10 * np.max(np.abs(refb['flattheory']))
    planb.close()


def test_taylor_state_from_reference_fit_on_gpu(gpu_required):
    """The 'taylor' generator (monomials evaluated inside the contraction kernel) on the state produced by the reference's own in-tree
    ``TaylorEmulatorEngine._fit_no_operation`` (tests/golden/taylor_state.npz, generator tests/golden/make_taylor_state.py): theory vectors ==
    the generating polynomial, element-wise 1e-10 (the reference layout of center / powers / derivatives is what the CUDA path consumes)."""
    import desilike_b200
    data = np.load(os.path.join(HERE, 'golden', 'taylor_state.npz'))
    params = [str(p) for p in data['params']]
    T = len(data['powers'])
    ny = int(np.prod(data['yshape']))
    theory = dict(kind='taylor', params=params, center=data['center'], powers=data['powers'], tables={'y': data['derivatives'].reshape(T, ny)})
    obs = dict(name='poly', theory=theory, terms=[dict(scale=1., factors=[], table='y')], nin=ny, wmatrix=None, offset=None, data=np.zeros(ny))
    spec = dict(params=params, priors={p: None for p in params}, solved=[], observables=[obs], precision=[np.eye(ny)])
    for reduce_rank in (True, False):
        plan = desilike_b200.Plan(spec, reduce_rank=reduce_rank)
        got = plan.theory(data['xtest'])
        want = data['ytest'].reshape(len(got), ny)
        assert np.max(np.abs(got - want) / np.abs(want)) < 1e-10
        chi2 = -2. * plan.eval(data['xtest'])['loglikelihood']
        assert np.allclose(chi2, np.sum(want**2, axis=-1), rtol=1e-11)
        plan.close()


@pytest.mark.parametrize('interp', ['scipy', 'interpax'])
def test_kaiser_merged_combine_and_interpolants(gpu_required, interp):
    """Direct Kaiser theory: (a) the tracer combine b1^2 dd + 2 b1 dt + tt formed INSIDE the Kaiser kernel (one merged output segment,
    lowering._merge_kaiser_terms) == the three-table plan == the oracle; (b) both interpolants of the reference's ``interp1d(method='cubic')``:
    SciPy not-a-knot (NumPy branch, jax.py:263-265) and the C1 local cubic of interpax (jax branch, jax.py:258-261; restated, oracle/core.py).
    Theory vectors element-wise to 1e-10 away from zero crossings."""
    import desilike_b200
    for spec in [configs.config_c1(), configs.config_kaiser_window(eftlike=True, solved=True), configs.config_kaiser_window(eftlike=False, solved=False, nmu=5)]:
        for obs in spec['observables']: obs['theory']['interp'] = interp
        X = configs.sample_points(spec, 150, seed=8)
        ref = pipeline.evaluate(spec, X, return_theory=True)
        for merge in (True, False):
            plan = desilike_b200.Plan(spec, merge_terms=merge)
            assert ('combo0' in plan.lowered.gens[0].tables) == merge
            theory, got = plan.theory(X), plan.eval(X)
            scale = np.max(np.abs(ref['flattheory']), axis=0)
            assert np.max(np.abs(theory - ref['flattheory']) / scale) < 1e-10
            away = np.abs(ref['flattheory']) > 1e-3 * scale
            assert np.max(np.abs(theory - ref['flattheory'])[away] / np.abs(ref['flattheory'])[away]) < 1e-10
            assert np.all(np.abs(got['logposterior'] - ref['logposterior']) <= 1e-8 * np.maximum(1., np.abs(ref['logposterior'])))
            plan.close()
    # the two interpolants differ at the level the jax / NumPy branches of the reference differ (well above 1e-10): the switch is not a no-op
    spec = configs.config_c1()
    X = configs.sample_points(spec, 20, seed=9)
    t0 = pipeline.evaluate(spec, X, return_theory=True)['flattheory']
    spec['observables'][0]['theory']['interp'] = 'interpax'
    t1 = pipeline.evaluate(spec, X, return_theory=True)['flattheory']
    assert 1e-9 < np.max(np.abs(t1 - t0) / np.max(np.abs(t0), axis=0)) < 1e-2


def test_kaiser_window_of_knots_edge_cases(gpu_required):
    """The Kaiser kernel builds the spline only on the window of template knots a point reaches: points at the corners of the prior box, theory
    nodes reaching both ends of the template grid (window = whole grid), and a prior without limits (window sized for the whole grid)."""
    import desilike_b200
    spec = configs.config_c1()
    X = configs.sample_points(spec, 64, seed=10)
    X[:4, 0] = [0.8, 1.2, 0.8, 1.2]; X[:4, 1] = [0.8, 0.8, 1.2, 1.2]     # corners of uniform[0.8, 1.2]^2
    ref = pipeline.evaluate(spec, X, return_theory=True)
    for priors in (spec['priors'], dict(spec['priors'], qpar=None, qper=None)):
        plan = desilike_b200.Plan(dict(spec, priors=priors))
        theory = plan.theory(X)
        assert np.max(np.abs(theory - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
        plan.close()
    spec = configs.config_kaiser_window(eftlike=False, solved=False)
    th = spec['observables'][0]['theory']
    th['k'] = np.geomspace(1.3e-3, 0.75, len(th['k']))     # nodes close to both ends of the template grid geomspace(1e-3, 1, 500)
    spec = configs._finish(dict(spec, solved=[]), [(spec['observables'][0]['kout'], th['ells'], 1e-4)], np.random.default_rng(1), configs._FID)
    X = configs.sample_points(spec, 40, seed=11)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    plan = desilike_b200.Plan(spec)
    theory = plan.theory(X)
    assert np.max(np.abs(theory - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    plan.close()


@pytest.mark.parametrize('interp_order', [1, 3])
def test_kaiser_correlation_function_matches_reference_class(gpu_required, interp_order):
    """KaiserTracerCorrelationFunctionMultipoles on the CUDA path (kaiser_kernel on kin + the P -> xi operator folded into the term matrices; linear
    and cubic resampling, reference base.py:127-136): theory vectors against the reference class's own outputs (tests/golden/kaiser_corr.npz) and
    the likelihood against the oracle on a larger seeded batch."""
    import os
    import desilike_b200
    from test_host_api import make_kaiser_corr_likelihood
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kaiser_corr.npz'))
    like, names = make_kaiser_corr_likelihood(interp_order, data)
    spec = like._finalize_spec(like.build_spec())
    plan = desilike_b200.Plan(spec)
    golden = data['corr{:d}'.format(interp_order)].reshape(len(data['X']), -1)
    got = plan.theory(data['X'])
    assert np.max(np.abs(got - golden)) < 1e-10 * np.max(np.abs(golden))     # relative to the largest entry of the vectors
    rng = np.random.default_rng(31)
    X = np.column_stack([rng.uniform(0.9, 1.1, 300), rng.uniform(0.9, 1.1, 300), rng.uniform(-0.1, 0.1, 300), rng.uniform(0.8, 1.2, 300), rng.uniform(1., 3., 300)])
    ref, res = pipeline.evaluate(spec, X), plan.eval(X)
    assert np.all(res['status'] == 0)
    assert np.allclose(res['loglikelihood'], ref['loglikelihood'], rtol=1e-9, atol=1e-7)
    plan.close()


@pytest.mark.parametrize('interp_order', [1, 3])
def test_eftlike_kaiser_correlation_function(gpu_required, interp_order):
    """EFTLikeKaiserTracerCorrelationFunctionMultipoles on the CUDA path: counter terms as feature-space weights before the folded P -> xi operator.
    Sampled counter terms: theory vectors against the reference class's outputs (tests/golden/kaiser_corr.npz: corr_eft*); analytically
    marginalised: log-posterior and marginalised values against the oracle."""
    import desilike_b200
    from test_host_api import make_kaiser_corr_likelihood
    data = np.load(os.path.join(HERE, 'golden', 'kaiser_corr.npz'))
    like, names = make_kaiser_corr_likelihood(interp_order, data, eftlike='sampled')
    plan = desilike_b200.Plan(like._finalize_spec(like.build_spec()))
    golden = data['corr_eft{:d}'.format(interp_order)].reshape(len(data['X']), -1)
    got = plan.theory(np.column_stack([data['X'], data['ct']]))
    assert np.max(np.abs(got - golden)) < 1e-10 * np.max(np.abs(golden))
    plan.close()
    like, names = make_kaiser_corr_likelihood(interp_order, data, eftlike='marg')
    spec = like._finalize_spec(like.build_spec())
    plan = desilike_b200.Plan(spec)
    rng = np.random.default_rng(37)
    X = np.column_stack([rng.uniform(0.9, 1.1, 200), rng.uniform(0.9, 1.1, 200), rng.uniform(-0.1, 0.1, 200), rng.uniform(0.8, 1.2, 200), rng.uniform(1., 3., 200)])
    ref, res = pipeline.evaluate(spec, X), plan.eval(X)
    assert np.all(res['status'] == 0)
    assert np.allclose(res['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.allclose(res['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-8 * np.max(np.abs(ref['x_solved'])))
    plan.close()
