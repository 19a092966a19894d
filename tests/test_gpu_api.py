"""The reference-facing host API evaluated through the C-ABI on the GPU; these read like the reference's own tests
(likelihoods/tests/test_galaxy_clustering.py:6-16, observables/tests/test_galaxy_clustering.py:128-174, tests/test_base.py:246-294).
Run with ``-m gpu``."""
import json
import os

import numpy as np
import pytest

import desilike_b200 as dl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import configs, pipeline, synthetic
from test_host_api import make_kaiser_likelihood

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def kaiser_likelihood(**kwargs):
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    theory = dl.KaiserTracerPowerSpectrumMultipoles(template=template)
    observable = dl.TracerPowerSpectrumMultipolesObservable(data={}, klim={0: [0.02, 0.2, 0.005], 2: [0.02, 0.2, 0.005], 4: [0.02, 0.2, 0.005]},
                                                            theory=theory, covariance=np.eye(108), **kwargs)
    return dl.ObservablesGaussianLikelihood(observables=[observable])


def test_likelihood_is_zero_at_data_generating_parameters(gpu_required):
    """reference likelihoods/tests/test_galaxy_clustering.py:6-16: data = theory at default params, identity covariance."""
    like = kaiser_likelihood()
    assert like.varied_params.names() == ['qpar', 'qper', 'dm', 'df', 'b1', 'sn0']
    assert abs(like()) < 1e-12
    assert like.flatdiff.shape == (108,) and np.max(np.abs(like.flatdiff)) < 1e-9 * np.max(np.abs(like.flattheory))
    assert like(b1=2.) < -1.
    assert np.allclose(like.flatdiff, like.flattheory - like.flatdata)
    assert abs(like.loglikelihood + like.logprior - like(b1=2.)) < 1e-9


def test_vmap_conventions(gpu_required):
    """reference base.py:162-226: result shaped params.shape; errors='return' -> (results, errs); return_derived -> tuple."""
    like = kaiser_likelihood()
    rng = np.random.default_rng(0)
    params = {'b1': rng.uniform(1., 2., (3, 4)), 'qpar': rng.uniform(0.95, 1.05, (3, 4))}
    res = dl.vmap(like, backend='cuda')(params)
    assert res.shape == (3, 4)
    single = np.array([[like(b1=params['b1'][i, j], qpar=params['qpar'][i, j]) for j in range(4)] for i in range(3)])
    assert np.allclose(res, single, rtol=1e-12, atol=1e-9)
    (res2, derived), errs = dl.vmap(like, backend='cuda', errors='return', return_derived=True)(params)
    assert errs == {} and np.array_equal(res2, res)
    assert derived['loglikelihood'].shape == (3, 4) and np.allclose(derived['loglikelihood'] + derived['logprior'], res)
    with pytest.raises(ValueError):
        dl.vmap(like)({'b1': 1.})
    with pytest.raises(ValueError):
        dl.vmap(like)({'b1': np.zeros(3), 'qpar': np.ones(4)})
    with pytest.raises(dl.PipelineError):
        dl.vmap(like)({'unknown': np.zeros(3)})
    out = dl.vmap(like)({'b1': np.array([1.5, 5.])})   # b1 prior is uniform [0, 4]
    assert np.isfinite(out[0]) and np.isneginf(out[1])


@pytest.mark.parametrize('name', ['c1_kaiser_shapefit', 'kaiser_window', 'eftlike_sampled', 'eftlike_marg', 'eftlike_best', 'eftlike_prec'])
def test_api_matches_reference_outputs(gpu_required, name):
    """Mirror classes -> plan -> CUDA vs the reference's own outputs on the same inputs."""
    data = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    case = json.loads(str(data['case']))
    spec = getattr(configs, case['config'])(**case['kwargs'])
    like = make_kaiser_likelihood(spec, eftlike=case['eftlike'], mode=case['mode'])
    X = data['X']
    params = {n: X[:, i] for i, n in enumerate(spec['params'])}
    (res, derived), errs = dl.vmap(like, errors='return', return_derived=True)(params)
    ref = data['logposterior']
    finite = np.isfinite(ref)
    assert errs == {} and np.array_equal(np.isneginf(ref), np.isneginf(res))
    assert np.allclose(res[finite], ref[finite], rtol=1e-9, atol=1e-7)
    tagged = bool(spec['solved']) and case['mode'] != '.prec'
    dll = derived['loglikelihood'][()] if tagged else derived['loglikelihood']
    assert np.allclose(dll[finite], data['loglikelihood'][finite], rtol=1e-9, atol=1e-7)
    if tagged:   # '.prec' parameters are marginalised into the precision: no solved values
        orc = pipeline.evaluate(spec, X, return_hessian=True)
        for i, s in enumerate(spec['solved']):
            assert np.allclose(derived[s['name']], orc['x_solved'][:, i], rtol=1e-7, atol=1e-7)
        # derivs-tagged Hessian entries (reference likelihoods/base.py:342-352, :388-390, :409-411): likelihood + prior Hessian = posterior Hessian
        names = [s['name'] for s in spec['solved']]
        assert derived['loglikelihood'].derivs[0] == () and derived['loglikelihood'].shape == X.shape[:1] + (1 + len(names) * (len(names) + 1) // 2,)
        for i, a in enumerate(names):
            for j, b in enumerate(names[i:], start=i):
                hp = derived['loglikelihood'][(a, b)] + derived['logprior'][(b, a)]
                assert np.allclose(hp[finite], orc['hessian'][finite, i, j], rtol=1e-7, atol=1e-9 * np.max(np.abs(orc['hessian'][finite])))
        prior = spec['solved'][0].get('prior') or {}
        if prior.get('dist') == 'norm':
            assert np.allclose(derived['logprior'][(names[0], names[0])], -1. / prior['scale']**2)
        assert np.allclose(derived['logprior'][()][finite], data['logprior'][finite], rtol=1e-9, atol=1e-7)


def test_marg_invariant_under_sum_and_sumlikelihood(gpu_required):
    """(L + L)() == 2 L() - logprior (reference observables/tests/test_galaxy_clustering.py:174) for independent copies."""
    spec = configs.config_kaiser_window(eftlike=False, solved=False)
    like1, like2 = make_kaiser_likelihood(spec), make_kaiser_likelihood(spec)
    both = dl.SumLikelihood([like1, like2])
    vals = dict(b1=1.8, qpar=1.02, sn0=0.05)
    total, single = both(**vals), like1(**vals)
    assert abs(total - (2. * single - like1.logprior)) < 1e-8 * max(1., abs(total))


def test_emulated_likelihood_from_file(gpu_required, tmp_path):
    """Emulator file -> EmulatedCalculator.load -> tracer theory -> observable -> likelihood on the GPU == oracle of the spec."""
    spec = configs.config_c2(nkin=30, nbins=8, order=2)
    th = spec['observables'][0]['theory']
    kin, ells = np.arange(1, 31) * (0.35 / 30), (0, 2, 4)
    tables = {name: th['tables'][name].reshape(-1, 3, 30) for name in ['pk_dd', 'pk_dt', 'pk_tt']}
    fn = str(tmp_path / 'emu.npy')
    dl.Emulator.from_taylor(th['params'], th['center'], th['powers'], tables, fixed=dict(k=kin, ells=ells, z=1.)).save(fn)
    theory = dl.KaiserTracerPowerSpectrumMultipoles(pt=dl.EmulatedCalculator.load(fn))
    obs0 = spec['observables'][0]
    kout = configs.klim_edges((0.02, 0.02 + 8 * 0.005, 0.005))[:8]
    observable = dl.TracerPowerSpectrumMultipolesObservable(data=obs0['data'], wmatrix=obs0['wmatrix'], kin=kin, ellsin=ells, k=[kout] * 3, ells=ells,
                                                            theory=theory, shotnoise=1e4)
    like = dl.ObservablesGaussianLikelihood([observable], precision=spec['precision'][0])
    for p in like.all_params:
        if p.name in spec['params']: p.update(fixed=False, prior=spec['priors'][p.name])
        else: p.update(fixed=True)
    X = configs.sample_points(spec, 100, seed=1)
    res = dl.vmap(like)({n: X[:, i] for i, n in enumerate(spec['params'])})
    assert np.allclose(res, pipeline.evaluate(spec, X)['logposterior'], rtol=1e-9, atol=1e-7)


def test_bao_correlation_function_likelihood(gpu_required):
    """Damped BAO xi_ell(s) with analytically marginalised broadband through the mirror classes."""
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    theory = dl.DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=template, ells=(0, 2))
    for p in theory.params.select(basename='al*'):
        p.update(derived='.marg' if p.basename.endswith(('_0', '_2')) else False, fixed=not p.basename.endswith(('_0', '_2')))
    s = np.arange(20., 152., 4.)
    observable = dl.TracerCorrelationFunctionMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, s=[s, s], ells=(0, 2), theory=theory,
                                                                  covariance=np.diag(np.full(66, 1e-8)))
    like = dl.ObservablesGaussianLikelihood([observable])
    assert like.solved_params.names() == ['al0_0', 'al0_2', 'al2_0', 'al2_2']
    built = like.spec
    rng = np.random.default_rng(2)
    X = np.column_stack([rng.uniform(0.97, 1.03, 40), rng.uniform(0.97, 1.03, 40), rng.uniform(1.8, 2.2, 40), rng.uniform(0.9, 1.1, 40), rng.uniform(2., 4., 40)])
    assert built['params'] == ['qpar', 'qper', 'b1', 'dbeta', 'sigmas']
    (res, derived) = dl.vmap(like, return_derived=True)({n: X[:, i] for i, n in enumerate(built['params'])})
    orc = pipeline.evaluate(built, X)
    assert np.max(np.abs(res - orc['logposterior']) / np.maximum(1., np.abs(orc['logposterior']))) < 1e-7
    assert np.allclose(derived['al0_0'], orc['x_solved'][:, 0], rtol=1e-6, atol=1e-9)


def test_fit_taylor_emulator_on_gpu(gpu_required, tmp_path):
    """``Emulator.fit_taylor``: Taylor coefficients of the Kaiser pt tables from GPU theory evaluations on a finite-difference stencil
    (the .fit half of the emulator API, reference emulators/__init__.py:438-507).  Pins: the coefficients equal the oracle's
    finite-difference fit of the oracle theory; the emulated theory reproduces the direct theory at the expansion centre and
    differs away from it by the truncation error only (reference emulators/tests/test_taylor.py:90-104); save -> load -> same output."""
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    k, ells = np.linspace(0.01, 0.25, 30), (0, 2, 4)
    direct = dl.KaiserTracerPowerSpectrumMultipoles(template=template, k=k, ells=ells)
    delta = [0.02, 0.02, 0.02, 0.04]
    emu = dl.Emulator.fit_taylor(direct, order=3, delta=delta)
    eng = emu.state['engines']['pk_dd']
    assert list(eng['params']) == ['qpar', 'qper', 'dm', 'df'] and np.allclose(eng['center'], [1., 1., 0., 1.])
    ref = configs._kaiser_taylor_engine(k, ells, 3, configs.shapefit_template(), delta=delta)
    assert np.array_equal(eng['powers'], ref['powers'])
    rng = np.random.default_rng(0)
    dx = rng.uniform(-1., 1., (64, 4)) * np.array(delta)
    mono = np.prod(dx[:, None, :]**ref['powers'][None, :, :], axis=-1)
    for name in ['pk_dd', 'pk_dt', 'pk_tt']:   # identically-zero derivatives are rounding noise: compare the polynomials, not the coefficients
        got = np.asarray(emu.state['engines'][name]['derivatives']).reshape(len(ref['powers']), -1)
        pg, pr = mono.dot(got), mono.dot(ref['tables'][name])
        assert np.max(np.abs(pg - pr)) < 1e-9 * np.max(np.abs(pr)), name
    fn = str(tmp_path / 'fitted.npy')
    emu.save(fn)
    emulated = dl.KaiserTracerPowerSpectrumMultipoles(pt=dl.EmulatedCalculator.load(fn))

    def likelihood(theory):
        observable = dl.TracerPowerSpectrumMultipolesObservable(data={'b1': 2.}, k=[k] * 3, ells=ells, theory=theory, covariance=np.eye(90) * 1e4)
        return dl.ObservablesGaussianLikelihood([observable])

    like_d, like_e = likelihood(direct), likelihood(emulated)
    like_d(b1=2.); like_e(b1=2.)
    scale = np.max(np.abs(like_d.flattheory))
    assert np.max(np.abs(like_e.flattheory - like_d.flattheory)) < 1e-9 * scale          # at the centre
    like_d(b1=2., qpar=1.03, dm=0.02); like_e(b1=2., qpar=1.03, dm=0.02)
    err = np.max(np.abs(like_e.flattheory - like_d.flattheory)) / scale
    assert 1e-12 < err < 5e-3                                                              # away: truncation error of an order-3 expansion


@pytest.mark.parametrize('marg,engine', [(False, 'taylor'), (True, 'taylor'), (True, 'mlp')])
def test_lpt_velocileptors_table_model(gpu_required, marg, engine):
    """Velocileptors-style table model behind a Taylor emulator of pktable (reference full_shape.py:1182-1186, :1295-1313) on the CUDA path:
    theory vectors vs the un-collapsed restatement (Taylor predict -> bias combine -> cubic re-interpolation, pinned against the reference's own
    functions by tests/golden/lpt_combine.npz), log-posterior and marginalised alpha / sn values vs the oracle; 6 theta-dependent
    Jacobian rows exercise the 8-slot contraction variant."""
    import lpt_helpers
    like, info = lpt_helpers.make_lpt_likelihood(marg=marg, engine=engine)
    plan, spec = like.plan, like.spec
    rng = np.random.default_rng(4)
    N = 500
    X = np.column_stack([rng.uniform(0.95, 1.05, N), rng.uniform(0.95, 1.05, N), rng.uniform(-0.05, 0.05, N), rng.uniform(0.9, 1.1, N),
                         rng.uniform(1.5, 2.2, N), rng.uniform(-1., 1., N), rng.uniform(-1., 1., N)]
                        + ([] if marg else [rng.uniform(-5., 5., N) for i in range(3)] + [rng.uniform(-0.5, 0.5, N) for i in range(3)]))
    direct = lpt_helpers.direct_theory(info, X[:50], spec['params'])
    assert np.max(np.abs(plan.theory(X[:50]) - direct)) < 1e-10 * np.max(np.abs(direct))
    ref = pipeline.evaluate(spec, X)
    got = plan.eval(X)
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    if marg:
        assert np.max(np.abs(got['x_solved'] - ref['x_solved']) / np.maximum(1., np.abs(ref['x_solved']))) < 1e-7
    # calculator-style call (reference base.py:124-128, likelihoods/base.py:245)
    point = dict(zip(spec['params'], X[0]))
    assert abs(like(**point) - ref['logposterior'][0]) < 1e-8 * max(1., abs(ref['logposterior'][0]))


def test_torch_custom_op(gpu_required):
    """torch.ops.desilike_b200.eval_batch on device tensors == Plan.eval (C-ABI on the current torch stream)."""
    import torch
    from desilike_b200 import torch_ops
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True)
    plan = dl.Plan(spec)
    pid = torch_ops.register_plan(plan)
    X = configs.sample_points(spec, 333, seed=9)
    ref = plan.eval(X)
    with torch.cuda.stream(torch.cuda.Stream()):
        lp, ll, pr, xs, st = torch.ops.desilike_b200.eval_batch(pid, torch.as_tensor(X, device='cuda'))
        torch.cuda.current_stream().synchronize()
    assert np.array_equal(lp.cpu().numpy(), ref['logposterior']) and np.array_equal(ll.cpu().numpy(), ref['loglikelihood'])
    assert np.array_equal(pr.cpu().numpy(), ref['logprior']) and np.array_equal(st.cpu().numpy(), ref['status'])
    assert np.array_equal(xs.cpu().numpy()[:, plan._unperm], ref['x_solved'])
    torch_ops.unregister_plan(pid)
    plan.close()


def test_systematic_templates(gpu_required):
    """Systematic templates of the windowed power spectrum (reference window.py:470-473, :1285-1287), one sampled and one analytically
    marginalised amplitude, through the mirror classes on the CUDA path vs the oracle."""
    from test_host_api import make_systematics_likelihood
    like, base_spec, templates, kout = make_systematics_likelihood()
    spec = like.spec
    X0 = configs.sample_points(base_spec, 200, seed=13)
    X = np.column_stack([X0, np.random.default_rng(1).uniform(-1., 1., 200)])
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = like.plan.eval(X)
    assert np.max(np.abs(like.plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
    assert np.max(np.abs(got['logposterior'] - ref['logposterior']) / np.maximum(1., np.abs(ref['logposterior']))) < 1e-8
    assert np.max(np.abs(got['x_solved'] - ref['x_solved'])) < 1e-7 * max(1., np.max(np.abs(ref['x_solved'])))


def test_vmap_error_conventions_on_failed_points(gpu_required):
    """reference base.py:193-226 on the CUDA path: a point whose normal equations are singular is an error -- errors='raise' raises,
    'return' reports it in ``errs`` and fills its slot with a copy of the first good result, 'nan' fills NaN; a non-finite theory is not an
    exception (NaN propagates, as in the reference where samplers map it to -inf, samplers/base.py:187-191)."""
    spec = configs.config_kaiser_window(eftlike=True, solved=True)
    like = make_kaiser_likelihood(spec, eftlike=True, mode='.marg')
    for p in like.all_params:
        if p.name == 'b1': p.update(prior=None)
    rng = np.random.default_rng(4)
    n = 6
    params = {'b1': rng.uniform(1.5, 2.5, n), 'qpar': rng.uniform(0.98, 1.02, n)}
    good = dl.vmap(like)(params)
    assert np.all(np.isfinite(good))
    params['b1'][2] = 1e200          # overflows the theory: NaN, not an error
    res, errs = dl.vmap(like, errors='return')(params)
    assert errs == {} and np.isnan(res[2]) and np.allclose(np.delete(res, 2), np.delete(good, 2))
    # singular normal equations at every point but handled per point: drop the prior of a solved parameter and zero its template
    plan = like.plan
    spec2 = dict(plan.spec)
    spec2['solved'] = [dict(s) for s in spec2['solved']]
    name = spec2['solved'][0]['name']
    spec2['solved'][0]['prior'] = None
    spec2['observables'] = [dict(o, terms=[dict(t, colvec=np.zeros_like(t['colvec'])) if t.get('solved') == name else t for t in o['terms']]) for o in spec2['observables']]

    class Wrapped(object):
        def __init__(self, like, plan): self._like, self.plan = like, plan
        def _points(self, params): return self._like._points(params)

    bad = Wrapped(like, dl.Plan(spec2))
    params['b1'][2] = 2.
    with pytest.raises(ArithmeticError):
        dl.vmap(bad, errors='raise')(params)
    res, errs = dl.vmap(bad, errors='return')(params)
    assert sorted(errs) == list(range(n)) and isinstance(errs[0][0], ArithmeticError) and np.all(np.isnan(res))   # no good point to copy from
    res = dl.vmap(bad, errors='nan')(params)
    assert np.all(np.isnan(res))


def test_derived_expression_parameter_reparametrisation(gpu_required):
    """The reference's own pin for derived-expression parameters (desilike/tests/test_base.py:156-165), through the mirror API on the CUDA path:
    ``likelihood.all_params['b1'].update(derived='{b}**2', prior=None); likelihood.all_params['b'] = {'prior': {'limits': [0., 2.]}}`` --
    the marginalised likelihood at b = sqrt(1.5) equals the one at b1 = 1.5.  Also: a later Parameter.update rebuilds the plan."""
    like = kaiser_likelihood()
    for p in like.all_params:
        if p.basename == 'sn0': p.update(derived='.marg')
    like(b1=1.5)
    bak = like.loglikelihood
    assert 'b1' in like.plan.params
    like.all_params['b1'].update(derived='{b}**2', prior=None)
    like.all_params['b'] = {'prior': {'limits': [0., 2.]}}
    assert 'b' in like.varied_params.names() and 'b1' not in like.varied_params.names()
    like(b=1.5**0.5)
    assert 'b' in like.plan.params and 'b1' not in like.plan.params       # the plan was rebuilt for the new parametrisation
    assert np.allclose(like.loglikelihood, bak, rtol=1e-12, atol=1e-9)
    bs = np.linspace(0.8, 1.6, 7)
    res = dl.vmap(like)({'b': bs})
    like.all_params['b1'].update(derived=False, prior=dict(limits=[0., 4.]))
    like._extra_params = dl.ParameterCollection()
    res1 = dl.vmap(like)({'b1': bs**2})
    assert np.allclose(res, res1, rtol=1e-12, atol=1e-9)
    assert np.isneginf(dl.vmap(like)({'b1': np.array([5.])}))[0]   # b1 prior limits [0, 4] are back


def test_prec_premarginalisation_couples_observables_of_one_likelihood(gpu_required):
    """'.prec' parameters entering TWO observables of one likelihood: the reference updates that likelihood's full precision (likelihoods/base.py:
    298-308), which couples the observables (one dense block); the same parameters shared between two likelihoods are an error (:285-288)."""
    base = configs.config_kaiser_window(eftlike=True, solved=True)
    obs0 = base['observables'][0]
    obs1 = dict(obs0, name='LRG2', data=obs0['data'] * 1.01)
    spec = dict(base, observables=[obs0, obs1], precision=[base['precision'][0], 0.8 * base['precision'][0]], prec_groups=[[0, 1]],
                solved=[dict(s, mode='prec') for s in base['solved']], fiducial={n: configs._FID[n] for n in base['params']})
    X = configs.sample_points(spec, 40, seed=12)
    ref = pipeline.evaluate(spec, X)
    plan = dl.Plan(spec)
    assert len(plan.spec['precision']) == 1 and plan.spec['precision'][0].shape == (2 * len(obs0['data']),) * 2   # one coupled block
    got = plan.eval(X)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)
    plan.close()
    # block-by-block treatment (the previous behaviour) gives a different answer: the coupling matters
    wrong = pipeline.evaluate(dict(spec, prec_groups=[[0, 1]], precision=[spec['precision'][0], spec['precision'][1]]), X)['logposterior']
    indep = [pipeline.marginalize_precision(dict(spec, observables=[o], precision=[p], prec_groups=[[0]])) for o, p in zip(spec['observables'], spec['precision'])]
    assert not np.allclose(indep[0]['precision'][0], plan.spec['precision'][0][:len(obs0['data']), :len(obs0['data'])], rtol=1e-6)
    assert np.allclose(wrong, ref['logposterior'])
    for groups in ([[0], [1]], None):
        bad = dict(spec)
        if groups is None: bad.pop('prec_groups')
        else: bad['prec_groups'] = groups
        with pytest.raises(ValueError):
            dl.Plan(bad)
        with pytest.raises(ValueError):
            pipeline.evaluate(bad, X)


def test_fit_taylor_and_qmc_samples_for_bao_theory(gpu_required, tmp_path):
    """Emulator construction beyond Kaiser: ``Emulator.fit_taylor`` of a damped-BAO P(k) tracer theory (theory
    evaluated on the GPU on the finite-difference stencil) -> emulator file -> ``DampedBAOWigglesTracerPowerSpectrumMultipoles(pt=...)`` ->
    the emulated likelihood reproduces the direct one around the expansion point; ``Emulator.qmc_samples``: the MLP training design
    (reference emulators/__init__.py:528-533) with the tables evaluated on the GPU == the oracle tables at the same points."""
    k = np.linspace(0.02, 0.3, 57)
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    theory = dl.DampedBAOWigglesTracerPowerSpectrumMultipoles(template=template, k=k, ells=(0, 2))
    for p in theory.params.select(basename='al*'): p.update(fixed=True)
    for name, value in [('b1', 2.), ('sigmas', 3.)]: theory.all_params[name].update(value=value)
    emu = dl.Emulator.fit_taylor(theory, order=4, delta=[0.01, 0.01, 0.05, 0.02, 0.3])
    fn = str(tmp_path / 'bao_emu.npy')
    emu.save(fn)
    pt = dl.EmulatedCalculator.load(fn)
    assert pt.emulator.state['engines']['power']['yshape'] == (2, 57)
    emulated = dl.DampedBAOWigglesTracerPowerSpectrumMultipoles(pt=pt, ells=(0, 2), kp=theory.kp)
    for p in emulated.params.select(basename='al*'): p.update(fixed=True)
    cov = np.diag(np.full(114, 4e4))
    likes = []
    for th in (theory, emulated):
        obs = dl.TracerPowerSpectrumMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, k=[k, k], ells=(0, 2), theory=th, covariance=cov)
        likes.append(dl.ObservablesGaussianLikelihood([obs]))
    assert likes[0].plan.params == likes[1].plan.params == ['qpar', 'qper', 'b1', 'dbeta', 'sigmas']
    rng = np.random.default_rng(3)
    params = {'qpar': rng.uniform(0.99, 1.01, 50), 'qper': rng.uniform(0.99, 1.01, 50), 'b1': rng.uniform(1.95, 2.05, 50),
              'dbeta': rng.uniform(0.98, 1.02, 50), 'sigmas': rng.uniform(2.8, 3.2, 50)}
    direct, emul = dl.vmap(likes[0])(params), dl.vmap(likes[1])(params)
    assert np.max(np.abs(emul - direct)) < 2e-3 * max(1., np.max(np.abs(direct)))    # 5th-order Taylor remainder inside the stencil
    # training design of an MLP emulator: QMC points, tables on the GPU
    samples = dl.Emulator.qmc_samples(theory, niterations=300, params=['qpar', 'qper', 'b1'])
    assert samples['X'].shape == (300, 3) and samples['Y']['power'].shape == (300, 2, 57)
    assert np.all(samples['X'] >= samples['lower']) and np.all(samples['X'] <= samples['upper'])
    tspec = dict(theory.theory_spec(), params=['qpar', 'qper', 'b1'])
    ref = pipeline.theory_tables(tspec, samples['X'][:40], ['qpar', 'qper', 'b1'])['power']
    got = samples['Y']['power'][:40].reshape(40, -1)
    assert np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=0)) < 1e-10


@pytest.mark.parametrize('marg', [False, True])
def test_lpt_physical_prior_basis(gpu_required, marg):
    """Velocileptors ``prior_basis='physical'`` (reference full_shape.py:1299-1307) on the CUDA path: theory vectors vs the direct restatement,
    log-posterior and the marginalised alpha*p / sn*p vs the oracle."""
    import lpt_helpers
    like, info = lpt_helpers.make_lpt_likelihood(marg=marg, physical=True)
    spec = like.spec
    rng = np.random.default_rng(6)
    N = 200
    X = np.column_stack([rng.uniform(0.97, 1.03, N) if n in ('qpar', 'qper', 'df') else rng.uniform(-0.03, 0.03, N) if n == 'dm' else
                         rng.uniform(1.2, 1.8, N) if n == 'b1p' else rng.normal(0., 0.3, N) for n in spec['params']])
    theory = like.plan.theory(X[:40])
    direct = lpt_helpers.direct_theory(info, X[:40], spec['params'])
    assert np.max(np.abs(theory - direct) / np.max(np.abs(direct), axis=0)) < 1e-10
    got, ref = like.plan.eval(X), pipeline.evaluate(spec, X)
    assert np.all(np.abs(got['logposterior'] - ref['logposterior']) <= 1e-8 * np.maximum(1., np.abs(ref['logposterior'])))
    if marg: assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-7)


def test_workload_generators_product_side(gpu_required):
    """bench.py builds its workloads from desilike_b200/workloads.py without the checker package: the synthetic data vector (theory at the fiducial point)
    then comes from the CUDA path itself and equals the oracle-built one to rounding; a subprocess confirms that no oracle module gets imported."""
    import subprocess, sys
    from desilike_b200 import workloads
    from oracle import configs, pipeline
    kw = dict(ntracers=2, nkin=40, nbins=10, hidden=(16, 16))
    ref = configs.config_c4(**kw)                      # oracle hooks (installed by oracle/configs.py)
    workloads.set_hooks(None, None)
    try: own = workloads.config_c4(**kw)               # default hooks: Plan.theory
    finally: workloads.set_hooks(theory_flat=pipeline.observable_flat, theory_tables=pipeline.theory_tables)
    for a, b in zip(own['observables'], ref['observables']):
        assert np.allclose(a['data'], b['data'], rtol=1e-11, atol=1e-11 * np.max(np.abs(b['data'])))
        assert np.array_equal(a['wmatrix'], b['wmatrix'])
    for pa, pb in zip(own['precision'], ref['precision']): assert np.allclose(pa, pb, rtol=1e-8)
    code = ("import sys, bench; spec, label = bench.build_spec('c5solved'); "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), [m for m in sys.modules if m.startswith('oracle')]; print(len(spec['params']))")
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]


def test_fit_taylor_kaiser_correlation_function(gpu_required):
    """``Emulator.fit_taylor`` of a Kaiser CORRELATION-FUNCTION theory emulates its ``pt`` = the P(k) tables on kin (reference full_shape.py:341-346);
    ``KaiserTracerCorrelationFunctionMultipoles(pt=...)`` folds the P -> xi operator into the Taylor coefficient matrices: the emulated likelihood
    reproduces the direct one at the expansion centre and within the truncation error around it."""
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    s, ells = np.arange(20., 152., 4.), (0, 2)
    direct = dl.KaiserTracerCorrelationFunctionMultipoles(template=template, s=s, ells=ells)
    emu = dl.Emulator.fit_taylor(direct, order=2, delta=[0.01, 0.01, 0.01, 0.02])
    assert emu.state['engines']['pk_dd']['yshape'] == (2, 300) and np.allclose(emu.state['fixed']['k'], direct.kin)
    emulated = dl.KaiserTracerCorrelationFunctionMultipoles(pt=emu.to_calculator(), s=s, ells=ells)

    def likelihood(theory):
        observable = dl.TracerCorrelationFunctionMultipolesObservable(data={'b1': 2.}, s=[s] * 2, ells=ells, theory=theory, covariance=np.eye(2 * s.size) * 1e-6)
        return dl.ObservablesGaussianLikelihood([observable])

    like_d, like_e = likelihood(direct), likelihood(emulated)
    like_d(b1=2.); like_e(b1=2.)
    scale = np.max(np.abs(like_d.flattheory))
    assert np.max(np.abs(like_e.flattheory - like_d.flattheory)) < 1e-9 * scale          # at the centre
    like_d(b1=1.7, qpar=1.004, dm=-0.003); like_e(b1=1.7, qpar=1.004, dm=-0.003)
    diff = np.max(np.abs(like_e.flattheory - like_d.flattheory))
    assert 0. < diff < 1e-4 * scale                                                        # truncation error of the order-2 expansion
