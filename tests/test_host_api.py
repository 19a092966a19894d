"""Host-side mirror of the reference interface (CPU part): the mirror classes produce a likelihood spec whose oracle
evaluation equals the outputs of the reference's own classes (golden fixtures); vmap argument validation; emulator file
round trip.  The GPU part (same objects evaluated through the C-ABI) is tests/test_gpu_api.py."""
import json
import os

import numpy as np
import pytest

import desilike_b200 as dl
from oracle import configs, pipeline, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


def make_kaiser_likelihood(spec, eftlike=False, mode=None):
    """Same construction as oracle/ref_harness.py:reference_kaiser_likelihood, with desilike_b200's mirror classes."""
    obs0 = spec['observables'][0]
    th = obs0['theory']
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    cls = dl.EFTLikeKaiserTracerPowerSpectrumMultipoles if eftlike else dl.KaiserTracerPowerSpectrumMultipoles
    theory = cls(template=template, mu=th['nmu'])
    nell = len(th['ells'])
    if obs0.get('wmatrix') is not None:
        kw = dict(wmatrix=np.asarray(obs0['wmatrix_raw']), kin=th['k'], ellsin=th['ells'], k=[obs0['kout']] * nell, ells=th['ells'], shotnoise=obs0.get('shotnoise'))
    else:
        kw = dict(k=[th['k']] * nell, ells=th['ells'])
    observable = dl.TracerPowerSpectrumMultipolesObservable(data=np.asarray(obs0['data']), theory=theory, **kw)
    likelihood = dl.ObservablesGaussianLikelihood(observables=[observable], precision=np.array(spec['precision'][0]))
    for param in likelihood.all_params:
        if param.name in spec['params']:
            param.update(fixed=False, prior=spec['priors'].get(param.name))
        elif any(s['name'] == param.name for s in spec.get('solved', [])):
            s = [s for s in spec['solved'] if s['name'] == param.name][0]
            param.update(fixed=False, value=s['value'], prior=s.get('prior'), derived=mode or '.marg')
        else:
            param.update(fixed=True)
    return likelihood


@pytest.mark.parametrize('name', ['c1_kaiser_shapefit', 'kaiser_window', 'eftlike_sampled', 'eftlike_marg', 'eftlike_best', 'eftlike_prec'])
def test_mirror_classes_reproduce_reference_outputs(name):
    data = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    case = json.loads(str(data['case']))
    spec = getattr(configs, case['config'])(**case['kwargs'])
    like = make_kaiser_likelihood(spec, eftlike=case['eftlike'], mode=case['mode'])
    built = like._finalize_spec(like.build_spec())
    assert built['params'] == spec['params']
    assert [s['name'] for s in built['solved']] == [s['name'] for s in spec['solved']]
    X = data['X']
    got = pipeline.evaluate(built, X, return_theory=True)
    assert np.max(np.abs(got['flattheory'] - data['flattheory']) / np.max(np.abs(data['flattheory']), axis=0)) < 1e-12
    finite = np.isfinite(data['logposterior'])
    assert np.allclose(got['logposterior'][finite], data['logposterior'][finite], rtol=1e-11, atol=2e-9)


def test_reference_default_parameters():
    """Names, defaults and priors follow the reference YAML (full_shape.yaml:43-83, power_template.yaml:279-338)."""
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, f_fid=0.8)
    theory = dl.KaiserTracerPowerSpectrumMultipoles(template=template)
    names = theory.all_params.names()
    assert names == ['qpar', 'qper', 'dm', 'dn', 'df', 'sigmapar', 'sigmaper', 'b1', 'sn0']
    assert theory.all_params['b1'].value == 1.5 and theory.all_params['dn'].fixed and theory.all_params['sigmapar'].fixed
    assert theory.varied_params.names() == ['qpar', 'qper', 'dm', 'df', 'b1', 'sn0']
    eft = dl.EFTLikeKaiserTracerPowerSpectrumMultipoles(template=template)
    assert [p.basename for p in eft.params] == ['b1', 'ct0_2', 'ct2_2', 'ct4_2', 'sn0', 'sn0_2', 'sn2_2', 'sn4_2']
    assert eft.params['ct0_2'].prior == dict(dist='norm', loc=0., scale=100.)


def test_vmap_argument_validation():
    """reference base.py:144-157: rank >= 1, equal shapes; :531-533 unknown parameter."""
    from desilike_b200.base import _check_params
    with pytest.raises(ValueError):
        _check_params({'a': 1.})
    with pytest.raises(ValueError):
        _check_params({'a': np.zeros(3), 'b': np.zeros(4)})
    params, shape = _check_params({'a': np.zeros((2, 3)), 'b': np.ones((2, 3))})
    assert shape == (2, 3) and params['a'].shape == (6,)
    with pytest.raises(ValueError):
        dl.vmap(None, backend='jax')
    spec = configs.config_c1()
    like = make_kaiser_likelihood(spec)
    like._plan = type('FakePlan', (), {'params': spec['params'], 'close': lambda self: None})()
    with pytest.raises(dl.PipelineError):
        like._points({'not_a_param': [1.]})
    X = like._points({'b1': [2., 3.]})
    assert X.shape == (2, 6) and np.array_equal(X[:, 4], [2., 3.]) and np.all(X[:, 0] == 1.)


def test_emulator_file_round_trip(tmp_path):
    """Emulator save -> load -> same state and same spec (reference emulators/tests/test_base.py:10-69)."""
    spec = configs.config_c2(nkin=30, nbins=8, order=2)
    th = spec['observables'][0]['theory']
    kin, ells = np.arange(1, 31) * (0.35 / 30), (0, 2, 4)
    tables = {name: th['tables'][name].reshape(-1, 3, 30) for name in ['pk_dd', 'pk_dt', 'pk_tt']}
    emu = dl.Emulator.from_taylor(th['params'], th['center'], th['powers'], tables, fixed=dict(k=kin, ells=ells, z=1.))
    fn = str(tmp_path / 'emu.npy')
    emu.save(fn)
    state = np.load(fn, allow_pickle=True)[()]
    for key in ['engines', 'xoperations', 'yoperations', 'defaults', 'fixed', 'varied_params', 'in_calculator_state', 'calculator__class__',
                'is_calculator_sequence', 'yaml_data', 'all_params']:
        assert key in state, key
    assert state['engines']['pk_dd']['name'] == 'taylor' and state['engines']['pk_dd']['yshape'] == (3, 30)
    pt = dl.EmulatedCalculator.load(fn)
    got = pt.theory_spec()
    assert got['kind'] == 'taylor' and np.array_equal(got['powers'], th['powers'])
    for name in ['pk_dd', 'pk_dt', 'pk_tt']:
        assert np.array_equal(got['tables'][name], th['tables'][name])
    assert np.array_equal(got['tables']['pk11_l0'], np.tile(th['tables']['pk_dd'][:, :30], (1, 3)))
    # emulated pt inside a tracer theory: same oracle result as the original spec
    theory = dl.KaiserTracerPowerSpectrumMultipoles(pt=pt)
    obs0 = spec['observables'][0]
    kout = configs.klim_edges((0.02, 0.02 + 8 * 0.005, 0.005))[:8]
    observable = dl.TracerPowerSpectrumMultipolesObservable(data=obs0['data'], wmatrix=obs0['wmatrix'], kin=kin, ellsin=ells, k=[kout] * 3, ells=ells,
                                                            theory=theory, shotnoise=1e4)
    like = dl.ObservablesGaussianLikelihood([observable], precision=spec['precision'][0])
    for p in like.all_params:
        if p.name in spec['params']: p.update(fixed=False, prior=spec['priors'][p.name])
        else: p.update(fixed=True)
    built = like._finalize_spec(like.build_spec())
    X = configs.sample_points(spec, 8, seed=1)
    cols = [spec['params'].index(n) for n in built['params']]
    assert np.allclose(pipeline.evaluate(built, X[:, cols])['logposterior'], pipeline.evaluate(spec, X)['logposterior'], rtol=1e-12, atol=1e-9)


def test_reference_fitted_taylor_state_through_emulator_file(tmp_path):
    """An emulator file whose Taylor engine arrays come from the reference's own fit (tests/golden/taylor_state.npz, produced by the unmodified
    ``TaylorEmulatorEngine._fit_no_operation``) and whose keys are the ones the reference writes: load -> theory spec keeps
    the reference layout bit for bit, and its prediction is the generating polynomial."""
    from oracle import core
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'taylor_state.npz'))
    params = [str(p) for p in data['params']]
    ny = tuple(int(i) for i in data['yshape'])
    engine = {'name': 'taylor', 'params': params, 'xshape': (len(params),), 'yshape': ny, 'xoperations': [], 'yoperations': [],
              'center': data['center'], 'derivatives': data['derivatives'], 'powers': data['powers']}
    state = {'engines': {'pk_dd': engine}, 'xoperations': [], 'yoperations': [], 'defaults': dict(zip(params, data['center'])),
             'fixed': {'k': np.linspace(0.01, 0.2, ny[1]), 'ells': (0, 2), 'z': 1., 'names': ['pk_dd']}, 'varied_params': params,
             'in_calculator_state': ['pk_dd'], 'calculator__class__': ('desilike.theories.galaxy_clustering.full_shape', 'KaiserPowerSpectrumMultipoles'),
             'is_calculator_sequence': False, 'yaml_data': {}, 'all_params': {'data': [dict(basename=p, namespace='', value=float(c), fixed=False, prior=None, ref=None, derived=False)
                                                                                       for p, c in zip(params, data['center'])]}}
    fn = str(tmp_path / 'ref_taylor.npy')
    np.save(fn, state, allow_pickle=True)
    got = dl.EmulatedCalculator.load(fn).theory_spec()
    assert got['kind'] == 'taylor' and list(got['params']) == params
    assert np.array_equal(got['powers'], data['powers']) and np.array_equal(got['center'], data['center'])
    assert np.array_equal(got['tables']['pk_dd'], data['derivatives'].reshape(len(data['powers']), -1))
    pred = core.taylor_predict(data['xtest'], got['center'], got['powers'], got['tables']['pk_dd'])
    assert np.max(np.abs(pred - data['ytest'].reshape(len(pred), -1))) < 1e-13 * np.max(np.abs(data['ytest']))


def test_mlp_emulator_file_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    kin = np.arange(1, 21) * (0.35 / 20)
    nin = 3 * 20
    mlp = configs.mlp_engine(rng, 3 * nin, hidden=(8, 8), kin=kin)
    tables = {'pk_dd': slice(0, nin), 'pk_dt': slice(nin, 2 * nin), 'pk_tt': slice(2 * nin, 3 * nin)}
    emu = dl.Emulator.from_mlp(['qpar', 'qper', 'dm', 'df'], mlp, tables, {n: (3, 20) for n in tables}, fixed=dict(k=kin, ells=(0, 2, 4), z=1.))
    fn = str(tmp_path / 'mlp.npy')
    emu.save(fn)
    got = dl.EmulatedCalculator.load(fn).theory_spec()
    from oracle import core
    X = rng.uniform(0.95, 1.05, (6, 4)); X[:, 2] -= 1.
    y0, y1 = core.mlp_predict(X, mlp), core.mlp_predict(X, got['mlp'])
    for name, sl in tables.items():
        assert np.allclose(y1[:, got['tables'][name]], y0[:, sl], rtol=1e-13)
    assert len(got['mlp']['layers']) == 3 and got['mlp']['layers'][0][0].shape == (4, 8)   # shared hidden stack kept once


def test_unknown_operation_fails_loudly(tmp_path):
    rng = np.random.default_rng(3)
    kin = np.arange(1, 11) * 0.03
    mlp = configs.mlp_engine(rng, 30, hidden=(4,), kin=kin)
    emu = dl.Emulator.from_mlp(['qpar', 'qper', 'dm', 'df'], mlp, {'pk_dd': slice(0, 30)}, {'pk_dd': (3, 10)}, fixed=dict(k=kin, ells=(0, 2, 4), z=1.))
    emu.state['engines']['pk_dd']['model_operations'][1]['expression'] = 'jnp.sin(v)'
    with pytest.raises(ValueError):
        emu.to_calculator().theory_spec()


def test_lpt_table_model_mlp_emulated():
    """The same table model behind an MLP emulator of pktable: table slices selected by index from the network output, re-interpolation as a linear
    map after the selection; lowered plan (NumPy mirror) == oracle pipeline == direct restatement (MLP predict -> combine -> interp1d)."""
    import lowered_numpy
    import lpt_helpers
    from test_capi_and_lowering import _mlp_features
    from desilike_b200.lowering import Lowered
    like, info = lpt_helpers.make_lpt_likelihood(marg=True, engine='mlp')
    spec = like._finalize_spec(like.build_spec())
    low = Lowered(spec)
    rng = np.random.default_rng(3)
    X = np.column_stack([rng.uniform(0.95, 1.05, 6), rng.uniform(0.95, 1.05, 6), rng.uniform(-0.05, 0.05, 6), rng.uniform(0.9, 1.1, 6),
                         rng.uniform(1.5, 2.2, 6), rng.uniform(-1., 1., 6), rng.uniform(-1., 1., 6)])
    direct = lpt_helpers.direct_theory(info, X, spec['params'])
    ref = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(ref['flattheory'] - direct)) < 1e-11 * np.max(np.abs(direct))
    theory = lowered_numpy.evaluate(low, X, feature_fn=_mlp_features, white=False)
    assert np.max(np.abs(theory - direct)) < 1e-11 * np.max(np.abs(direct))
    got = lowered_numpy.evaluate(low, X, feature_fn=_mlp_features)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize('marg', [False, True])
def test_lpt_velocileptors_table_model_lowering(marg):
    """Velocileptors-style table model (reference full_shape.py:1182-1186, :1295-1313) behind a Taylor emulator of pktable: the lowered plan
    (19 bias monomials x table maps with the cubic re-interpolation folded in; alpha / sn as marginalised Jacobian rows) evaluated by the
    NumPy mirror of the kernels reproduces (a) the un-collapsed restatement Taylor predict -> combine -> interp1d and (b) the oracle pipeline."""
    import lowered_numpy
    import lpt_helpers
    from desilike_b200.lowering import Lowered
    like, info = lpt_helpers.make_lpt_likelihood(marg=marg)
    spec = like._finalize_spec(like.build_spec())
    assert spec['params'] == (['qpar', 'qper', 'dm', 'df', 'b1', 'b2', 'bs'] if marg else ['qpar', 'qper', 'dm', 'df', 'b1', 'b2', 'bs', 'alpha0', 'alpha2', 'alpha4', 'sn0', 'sn2', 'sn4'])
    assert len(spec['solved']) == (6 if marg else 0)
    low = Lowered(spec)
    assert low.blocks[0]['nslots'] == (7 if marg else 1)
    rng = np.random.default_rng(2)
    X = np.column_stack([rng.uniform(0.95, 1.05, 9), rng.uniform(0.95, 1.05, 9), rng.uniform(-0.05, 0.05, 9), rng.uniform(0.9, 1.1, 9),
                         rng.uniform(1.5, 2.2, 9), rng.uniform(-1., 1., 9), rng.uniform(-1., 1., 9)] + ([] if marg else [rng.uniform(-5., 5., 9) for i in range(3)] + [rng.uniform(-0.5, 0.5, 9) for i in range(3)]))
    theory = lowered_numpy.evaluate(low, X, white=False)
    direct = lpt_helpers.direct_theory(info, X, spec['params'])
    assert np.max(np.abs(theory - direct)) < 1e-11 * np.max(np.abs(direct))
    ref = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(ref['flattheory'] - direct)) < 1e-11 * np.max(np.abs(direct))
    got = lowered_numpy.evaluate(low, X)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)
    if marg: assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-7)


def make_systematics_likelihood():
    """Kaiser + window likelihood with two systematic templates (one sampled amplitude, one analytically marginalised)."""
    spec = configs.config_kaiser_window(eftlike=False, solved=False)
    obs0 = spec['observables'][0]
    th = obs0['theory']
    nell, kout = len(th['ells']), obs0['kout']
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    theory = dl.KaiserTracerPowerSpectrumMultipoles(template=template, mu=th['nmu'])
    templates = {'syst_a': lambda ell, k: (ell == 0) * 3e3 * np.exp(-k / 0.05), 'syst_b': np.concatenate([1e3 * (kout / 0.1)**-1 * (ell == 2) for ell in th['ells']])}
    observable = dl.TracerPowerSpectrumMultipolesObservable(data=np.asarray(obs0['data']), theory=theory, wmatrix=np.asarray(obs0['wmatrix_raw']), kin=th['k'], ellsin=th['ells'],
                                                            k=[kout] * nell, ells=th['ells'], shotnoise=obs0.get('shotnoise'), systematic_templates=templates)
    like = dl.ObservablesGaussianLikelihood(observables=[observable], precision=np.array(spec['precision'][0]))
    for param in like.all_params:
        if param.name in spec['params']: param.update(fixed=False, prior=spec['priors'].get(param.name))
        elif param.name == 'syst_a': param.update(fixed=False, prior=dict(limits=[-2., 2.]))
        elif param.name == 'syst_b': param.update(fixed=False, prior=dict(dist='norm', loc=0., scale=3.), derived='.marg')
        else: param.update(fixed=True)
    return like, spec, templates, kout


def test_systematic_templates_lowering():
    """``flat += sum_i syst_i template_i`` after the window (reference window.py:470-473, :1285-1287): spec terms with an output-space vector;
    lowered plan (NumPy mirror) == oracle == the windowed theory without templates + the templates."""
    import lowered_numpy
    from desilike_b200.lowering import Lowered
    like, base_spec, templates, kout = make_systematics_likelihood()
    spec = like._finalize_spec(like.build_spec())
    assert spec['params'] == base_spec['params'] + ['syst_a'] and [s['name'] for s in spec['solved']] == ['syst_b']
    X0 = configs.sample_points(base_spec, 7, seed=12)
    a = np.linspace(-1., 1., 7)
    X = np.column_stack([X0, a])
    ref = pipeline.evaluate(spec, X, return_theory=True)
    base = pipeline.evaluate(base_spec, X0, return_theory=True)
    ta = np.concatenate([templates['syst_a'](ell, kout) for ell in (0, 2, 4)])
    assert np.allclose(ref['flattheory'], base['flattheory'] + a[:, None] * ta, rtol=1e-13, atol=1e-9)   # syst_b at its default 0
    low = Lowered(spec)
    assert len(low.blocks[0]['const_solved']) == 1
    got = lowered_numpy.evaluate(low, X, feature_fn=lowered_numpy.theory_features)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-9)


def test_correlation_window_constants_fold():
    """``WindowedCorrelationFunctionMultipoles._apply`` (reference window.py:717-728): matrix_diag (per-s multipole mixing), matrix_full, offset and
    the s mask fold into ONE matrix + offset of the observable spec; checked against the outputs of the reference's own method
    (tests/golden/xi_window.npz, generated by tests/golden/make_golden.py:xi_window)."""
    data = np.load(os.path.join(HERE, 'golden', 'xi_window.npz'))
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    theory = dl.DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=template, ells=(0, 2))
    sin, s = data['sin'], [data['s0'], data['s1']]
    nout = sum(len(ss) for ss in s)
    obs = dl.TracerCorrelationFunctionMultipolesObservable(data=np.zeros(nout), s=s, ells=(0, 2), theory=theory, matrix_diag=data['matrix_diag'], offset=data['offset_diag'])
    ospec = obs.observable_spec()
    for th, ref in zip(data['theory'], data['out_diag']):
        assert np.allclose(ospec['wmatrix'] @ th.ravel() + ospec['offset'], ref, rtol=1e-14, atol=1e-14)
    obs = dl.TracerCorrelationFunctionMultipolesObservable(data=np.zeros(nout), s=s, ells=(0, 2), theory=theory, wmatrix=data['matrix_full'], sin=sin,
                                                           matrix_diag=data['matrix_diag'], offset=data['offset_full'])
    ospec = obs.observable_spec()
    for th, ref in zip(data['theory'], data['out_full']):
        assert np.allclose(ospec['wmatrix'] @ th.ravel() + ospec['offset'], ref, rtol=1e-13, atol=1e-13)



def test_generic_table_bias_model():
    """TableBiasTracerPowerSpectrumMultipoles with a user-defined monomial list (here: b1^2, b1, 1 and a linear counter-term with scale 0.5)
    equals the hand-written combination of the emulated table slices."""
    import lpt_helpers
    like, info = lpt_helpers.make_lpt_likelihood(marg=False)
    emu_calc = like.observables[0].theory.pt
    params = [dl.Parameter('b1', value=1.5, prior=dict(limits=[0., 4.])), dl.Parameter('c0', value=0., prior=dict(dist='norm', loc=0., scale=10.))]
    monos = [('b1', 'b1'), ('b1',), ()] + [(0.5, ('c0',))] + [(0., ())] * 15
    theory = dl.TableBiasTracerPowerSpectrumMultipoles(pt=emu_calc, monomials=monos, params=params, k=info['k'], ells=info['ells'])
    terms = theory.terms()
    assert terms[0]['factors'] == [('b1', 2)] and terms[3]['scale'] == 0.5 and terms[3]['factors'] == [('c0', 1)]
    tspec = theory.theory_spec()
    names = ['qpar', 'qper', 'dm', 'df']
    X = np.array([[1.02, 0.98, 0.03, 1.05]])
    tabs = pipeline.theory_tables(tspec, X, names)
    from oracle import core
    diffs = X[0] - info['center']
    pktable = np.tensordot(np.prod(diffs**info['powers'], axis=-1), info['deriv'], axes=(0, 0))
    b1, c0 = 1.7, 3.
    direct = core.interp1d_cubic(info['k'], info['kpt'], (b1**2 * pktable[..., 0] + b1 * pktable[..., 1] + pktable[..., 2] + 0.5 * c0 * pktable[..., 3]).T).T.ravel()
    got = b1**2 * tabs['pktable_0'][0] + b1 * tabs['pktable_1'][0] + tabs['pktable_2'][0] + 0.5 * c0 * tabs['pktable_3'][0]
    assert np.allclose(got, direct, rtol=1e-12, atol=1e-8)


def test_covariance_corrections_match_reference():
    """ObservablesGaussianLikelihood from a covariance with scale_covariance and the Hartlap 2007 / Percival 2014 factors (reference
    likelihoods/base.py:605-656): factors, precision and log-posterior against the reference's own run (tests/golden/covariance_corrections.npz)."""
    data = np.load(os.path.join(HERE, 'golden', 'covariance_corrections.npz'))
    case = json.loads(str(data['case']))
    spec = configs.config_kaiser_window(eftlike=True, solved=True)
    like = make_kaiser_likelihood(spec, eftlike=True, mode='.marg')
    obs = like.observables[0]
    like = dl.ObservablesGaussianLikelihood(observables=[obs], covariance=data['covariance'], scale_covariance=case['scale_covariance'],
                                            correct_covariance={'nobs': case['nobs'], 'correction': 'hartlap-percival2014'})
    built = like._finalize_spec(like.build_spec())
    assert np.isclose(like.hartlap2007_factor, float(data['hartlap2007_factor']), rtol=1e-14)
    assert np.isclose(like.percival2014_factor, float(data['percival2014_factor']), rtol=1e-14)
    P = built['precision'][0] if isinstance(built['precision'], list) else built['precision']
    assert np.allclose(P, data['precision'], rtol=1e-9, atol=1e-9 * np.max(np.abs(data['precision'])))
    got = pipeline.evaluate(built, data['X'])
    for key in ['logposterior', 'loglikelihood', 'logprior']:
        ref = data[key]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isneginf(ref), np.isneginf(got[key]))
        assert np.allclose(got[key][fin], ref[fin], rtol=1e-10, atol=1e-8), key


def test_bao_broadband_terms_match_reference_classes():
    """Broadband terms of the damped-BAO tracers (reference bao.py:467-476, :527-534: al{ell}_{pow} (k / kp)^pow; :849-853, :905: (s / sp)^pow) through
    the mirror classes' spec (oracle evaluation) against the outputs of the reference's own classes (tests/golden/bao_broadband.npz)."""
    data = np.load(os.path.join(HERE, 'golden', 'bao_broadband.npz'))
    case = json.loads(str(data['case']))
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper']
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    for which, cls, x, bb, A in [('power', dl.DampedBAOWigglesTracerPowerSpectrumMultipoles, data['k'], case['bb_pk'], data['A_pk']),
                                 ('corr', dl.DampedBAOWigglesTracerCorrelationFunctionMultipoles, data['s'], case['bb_xi'], data['A_xi'])]:
        theory = cls(template=template, ells=(0, 2), **({'k': x} if which == 'power' else {'s': x}))
        assert np.isclose(theory.kp if which == 'power' else theory.sp, float(data['kp' if which == 'power' else 'sp']), rtol=1e-14)
        for p in theory.all_params: p.update(fixed=p.basename not in names + bb)
        params = names + bb
        terms = []
        for term in theory.terms():   # fixed broadband amplitudes (value 0) drop out, as in ObservablesGaussianLikelihood.build_spec
            factors = [(n, pw) for n, pw in term.get('factors', []) if n in params]
            if len(factors) == len(term.get('factors', [])): terms.append(dict(term, factors=factors))
        obs = dict(theory=dict(theory.theory_spec(), params=names), terms=terms, nin=theory.nin, wmatrix=None, offset=None)
        X = np.column_stack([data['X'], A])
        got = pipeline.observable_flat(obs, X, params, {})
        ref = data[which].reshape(len(X), -1)
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref)), which


def test_fiber_collision_kernels_match_reference_class():
    """Top-hat fiber-collision kernels on xi multipoles against the reference's own class (tests/golden/fiber_collisions.npz; reference
    window.py:1192-1249) and their folding into the observable spec (reference window.py:686-702, :717-728)."""
    from desilike_b200.observables import tophat_fiber_collisions_correlation
    data = np.load(os.path.join(HERE, 'golden', 'fiber_collisions.npz'))
    case = json.loads(str(data['case']))
    for n in 'ab':
        corr, unc = tophat_fiber_collisions_correlation(data['s'], tuple(data['ells']), ellsin=tuple(data['ellsin']), **case[n])
        assert np.allclose(corr, data['correlated_' + n], rtol=0., atol=1e-14) and np.allclose(unc, data['uncorrelated_' + n], rtol=0., atol=1e-14)
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    theory = dl.DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=template, ells=(0, 2))
    s = data['s'][6:]
    obs = dl.TracerCorrelationFunctionMultipolesObservable(data=np.zeros(2 * s.size), s=[s, s], ells=(0, 2), theory=theory, fiber_collisions=dict(fs=0.6, Dfc=3.))
    ospec = obs.observable_spec()
    th = np.random.default_rng(0).standard_normal((2, s.size))
    corr, unc = tophat_fiber_collisions_correlation(s, (0, 2), fs=0.6, Dfc=3.)
    assert np.allclose(ospec['wmatrix'] @ th.ravel() + ospec['offset'], (np.sum(corr * th[None], axis=1) + unc).ravel(), rtol=1e-14, atol=1e-14)


def test_power_fiber_collision_kernels_match_reference_class():
    """Top-hat fiber-collision kernels on P_ell(k) (Hahn et al. 2016) against the reference's own class (tests/golden/fiber_collisions.npz;
    reference window.py:967-1049) and their folding into the window matrix / offset of the observable (reference window.py:428-438)."""
    from desilike_b200.observables import tophat_fiber_collisions_power
    data = np.load(os.path.join(HERE, 'golden', 'fiber_collisions.npz'))
    k = data['p_k']
    corr, unc = tophat_fiber_collisions_power(k, (0, 2, 4), fs=0.5, Dfc=0.8)
    assert np.allclose(corr, data['p_correlated'], rtol=1e-13, atol=1e-15)
    assert np.allclose(unc, data['p_uncorrelated'], rtol=1e-12, atol=1e-12 * np.max(np.abs(data['p_uncorrelated'])))
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    rng = np.random.default_rng(3)
    power = rng.standard_normal((3, k.size))
    ref = np.sum(corr * power[None, :, None, :], axis=(1, 3)) + unc   # reference BaseFiberCollisionsPowerSpectrumMultipoles.calculate (window.py:812-821)
    # without a window matrix
    theory = dl.KaiserTracerPowerSpectrumMultipoles(template=template)
    obs = dl.TracerPowerSpectrumMultipolesObservable(data=np.zeros(3 * k.size), k=[k] * 3, ells=(0, 2, 4), theory=theory, fiber_collisions=dict(fs=0.5, Dfc=0.8))
    ospec = obs.observable_spec()
    assert np.allclose(ospec['wmatrix'] @ power.ravel() + ospec['offset'], ref.ravel(), rtol=1e-12, atol=1e-9)
    # behind a window matrix: W . (kernels . P + uncorrelated)
    kout = k[5:40:5]
    W = rng.standard_normal((3 * kout.size, 3 * k.size))
    theory = dl.KaiserTracerPowerSpectrumMultipoles(template=template)
    obs = dl.TracerPowerSpectrumMultipolesObservable(data=np.zeros(3 * kout.size), k=[kout] * 3, ells=(0, 2, 4), theory=theory, wmatrix=W, kin=k, ellsin=(0, 2, 4),
                                                     fiber_collisions=dict(fs=0.5, Dfc=0.8))
    ospec = obs.observable_spec()
    assert np.allclose(ospec['wmatrix'] @ power.ravel() + ospec['offset'], W @ ref.ravel(), rtol=1e-11, atol=1e-8)


def test_lpt_physical_prior_basis_lowering():
    """Velocileptors ``prior_basis='physical'`` (the reference default, full_shape.py:1275, :1299-1307) on the lowered plan: sigma8 / fsigma8 from
    scalar Taylor engines become derived-expression parameters, b1 .. b3 derived from them, the alpha / sn slices expanded in (b1, f); the NumPy
    mirror of the kernels == the direct restatement (Taylor predict -> physical mapping -> bias combine -> interp1d), sampled and marginalised."""
    import lowered_numpy
    import lpt_helpers
    from desilike_b200.lowering import Lowered
    for marg in (False, True):
        like, info = lpt_helpers.make_lpt_likelihood(marg=marg, physical=True)
        spec = like._finalize_spec(like.build_spec())
        assert [dp['name'] for dp in spec['derived']][:2] == ['sigma8', 'f'] and 'b1p' in spec['params'] and 'b1' not in spec['params']
        low = Lowered(spec, with_plain=True)
        rng = np.random.default_rng(5)
        X = np.column_stack([rng.uniform(0.97, 1.03, 12) if n in ('qpar', 'qper', 'df') else rng.uniform(-0.03, 0.03, 12) if n == 'dm' else
                             rng.uniform(1.2, 1.8, 12) if n == 'b1p' else rng.normal(0., 0.3, 12) for n in spec['params']])
        theory = lowered_numpy.evaluate(low, X, white=False)
        direct = lpt_helpers.direct_theory(info, X, spec['params'])
        assert np.max(np.abs(theory - direct) / np.max(np.abs(direct), axis=0)) < 1e-11
        got = lowered_numpy.evaluate(low, X)
        ref = pipeline.evaluate(spec, X)
        assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-10, atol=1e-8)
        if marg: assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-8, atol=1e-8)


def make_kaiser_corr_likelihood(interp_order, data, eftlike=None):
    """Mirror-class likelihood of the tests/golden/kaiser_corr.npz case (reference KaiserTracerCorrelationFunctionMultipoles + ShapeFit template).
    ``eftlike``: 'sampled' / 'marg' -> EFTLikeKaiserTracerCorrelationFunctionMultipoles with ct0_2, ct2_2 sampled / analytically marginalised."""
    s, names = data['s'], json.loads(str(data['case']))['names']
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, a=float(data['a']), kp=float(data['kp']))
    cls = dl.EFTLikeKaiserTracerCorrelationFunctionMultipoles if eftlike else dl.KaiserTracerCorrelationFunctionMultipoles
    theory = cls(template=template, ells=(0, 2), interp_order=interp_order)
    rng = np.random.default_rng(5)
    obs = dl.TracerCorrelationFunctionMultipolesObservable(data=rng.normal(0., 0.05, 2 * s.size) if eftlike == 'marg' else np.zeros(2 * s.size), s=[s, s], ells=(0, 2), theory=theory)
    like = dl.ObservablesGaussianLikelihood(observables=[obs], precision=np.eye(2 * s.size) * 1e4)
    if eftlike == 'sampled': names = names + ['ct0_2', 'ct2_2']
    for param in like.all_params:
        if param.name in names: param.update(fixed=False, prior=None)
        elif eftlike == 'marg' and param.name in ('ct0_2', 'ct2_2'): param.update(fixed=False, derived='.marg', prior=dict(dist='norm', loc=0., scale=50.))
        else: param.update(fixed=True)
    return like, names


@pytest.mark.parametrize('interp_order', [1, 3])
def test_kaiser_correlation_function_fold(interp_order):
    """KaiserTracerCorrelationFunctionMultipoles (reference full_shape.py:336-365, :553-574): Kaiser tables on kin, the P -> xi operator (linear or
    cubic resampling, base.py:127-136) folded into the term matrices.  Lowered plan (NumPy mirror of the kernels) == oracle == the reference
    class's own outputs (tests/golden/kaiser_corr.npz, generated by tests/golden/make_golden.py:kaiser_corr)."""
    import lowered_numpy
    from desilike_b200.lowering import Lowered
    data = np.load(os.path.join(HERE, 'golden', 'kaiser_corr.npz'))
    like, names = make_kaiser_corr_likelihood(interp_order, data)
    spec = like._finalize_spec(like.build_spec())
    assert spec['params'] == names
    th = spec['observables'][0]['theory']
    assert np.allclose(th['kin'], data['kin{:d}'.format(interp_order)], rtol=1e-14) and np.allclose(th['template']['k'], data['k_template'], rtol=1e-14)
    X, golden = data['X'], data['corr{:d}'.format(interp_order)].reshape(len(data['X']), -1)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    scale = np.max(np.abs(golden))
    assert np.max(np.abs(ref['flattheory'] - golden)) < 1e-13 * scale
    low = Lowered(spec, with_plain=True)
    got = lowered_numpy.evaluate(low, X, feature_fn=lowered_numpy.theory_features)
    assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-9, atol=1e-7)
    with pytest.raises(ValueError):
        dl.KaiserTracerCorrelationFunctionMultipoles(template=None, interp_order=2)


@pytest.mark.parametrize('interp_order', [1, 3])
def test_eftlike_kaiser_correlation_function_fold(interp_order):
    """EFTLikeKaiserTracerCorrelationFunctionMultipoles (reference full_shape.py:577-634, :664-684): the counter terms ct{ell}_2 (k / kp)^2 pk11_{ell=0}
    weight P(kin) BEFORE the folded P -> xi operator (term key ``featvec``).  Sampled: lowered plan == oracle == the reference class's outputs
    (tests/golden/kaiser_corr.npz: corr_eft*); analytically marginalised: lowered plan == oracle (x_solved too)."""
    import lowered_numpy
    from desilike_b200.lowering import Lowered
    data = np.load(os.path.join(HERE, 'golden', 'kaiser_corr.npz'))
    like, names = make_kaiser_corr_likelihood(interp_order, data, eftlike='sampled')
    spec = like._finalize_spec(like.build_spec())
    assert spec['params'] == names and names[-2:] == ['ct0_2', 'ct2_2']
    X, golden = np.column_stack([data['X'], data['ct']]), data['corr_eft{:d}'.format(interp_order)].reshape(len(data['X']), -1)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(ref['flattheory'] - golden)) < 1e-13 * np.max(np.abs(golden))
    got = lowered_numpy.evaluate(Lowered(spec, with_plain=True), X, feature_fn=lowered_numpy.theory_features)
    assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-9, atol=1e-7)
    like, names = make_kaiser_corr_likelihood(interp_order, data, eftlike='marg')
    spec = like._finalize_spec(like.build_spec())
    assert [s_['name'] for s_ in spec['solved']] == ['ct0_2', 'ct2_2']
    ref = pipeline.evaluate(spec, data['X'])
    got = lowered_numpy.evaluate(Lowered(spec), data['X'], feature_fn=lowered_numpy.theory_features)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('eftlike', [False, True])
def test_kaiser_correlation_function_with_emulated_pt(eftlike):
    """``KaiserTracerCorrelationFunctionMultipoles(pt=EmulatedCalculator)`` (and the EFT-like class): the reference emulates ``theory.pt`` = the Kaiser
    tables on kin (full_shape.py:341-346); here the P -> xi operator (and the counter-term weights on P(kin)) are folded into the Taylor coefficient
    matrices.  A Taylor emulator of the ORACLE's Kaiser tables on kin: at the expansion point the emulated xi equals the oracle's direct xi (tables
    through oracle get_corr -- independent of the product's operator) to round-off, nearby to the truncation error; the lowered plan (NumPy mirror)
    equals the oracle on the emulated spec."""
    import lowered_numpy
    from oracle import core
    from desilike_b200.lowering import Lowered
    data = np.load(os.path.join(HERE, 'golden', 'kaiser_corr.npz'))
    like, names = make_kaiser_corr_likelihood(1, data, eftlike='sampled' if eftlike else None)
    direct = like._finalize_spec(like.build_spec())
    th = direct['observables'][0]['theory']
    pnames = ['qpar', 'qper', 'dm', 'df']
    center, delta = np.array([1., 1., 0., 1.]), np.array([0.01, 0.01, 0.01, 0.02])
    ptheory = {k: v for k, v in th.items() if k not in ('s', 'kin', 'interp_order')}   # the P(k) theory on kin
    ptheory['params'] = pnames

    def func(P):
        tabs = pipeline.theory_tables(ptheory, P, pnames)
        return np.concatenate([tabs[n] for n in ('pk_dd', 'pk_dt', 'pk_tt')], axis=-1)
    powers, derivs = core.taylor_fit_polynomial(func, center, 4, 2, delta)
    nin = 2 * th['kin'].size
    tables = {n: derivs[:, i * nin:(i + 1) * nin].reshape(-1, 2, th['kin'].size) for i, n in enumerate(('pk_dd', 'pk_dt', 'pk_tt'))}
    pt = dl.Emulator.from_taylor(pnames, center, powers, tables, fixed=dict(k=th['kin'], ells=(0, 2), z=1.)).to_calculator()
    cls = dl.EFTLikeKaiserTracerCorrelationFunctionMultipoles if eftlike else dl.KaiserTracerCorrelationFunctionMultipoles
    obs = dl.TracerCorrelationFunctionMultipolesObservable(data=np.zeros(2 * data['s'].size), s=[data['s']] * 2, ells=(0, 2), theory=cls(pt=pt, ells=(0, 2)))
    elike = dl.ObservablesGaussianLikelihood(observables=[obs], precision=np.eye(2 * data['s'].size) * 1e4)
    for param in elike.all_params:
        param.update(fixed=param.name not in names, **({} if param.name not in names else dict(prior=None)))
    spec = elike._finalize_spec(elike.build_spec())
    assert spec['params'] == names and spec['observables'][0]['theory']['kind'] == 'taylor'
    extra = [5., -3.] if eftlike else []
    X0 = np.array([list(center) + [1.8] + extra])
    a, b = pipeline.evaluate(spec, X0, return_theory=True)['flattheory'], pipeline.evaluate(direct, X0, return_theory=True)['flattheory']
    assert np.max(np.abs(a - b)) < 1e-11 * np.max(np.abs(b))
    rng = np.random.default_rng(3)
    X = X0 + rng.uniform(-1., 1., (6, X0.shape[1])) * np.array(list(0.3 * delta) + [0.5] + [3.] * len(extra))
    a, b = pipeline.evaluate(spec, X, return_theory=True), pipeline.evaluate(direct, X, return_theory=True)
    assert np.max(np.abs(a['flattheory'] - b['flattheory'])) < 1e-4 * np.max(np.abs(b['flattheory']))   # truncation error of the order-2 expansion
    got = lowered_numpy.evaluate(Lowered(spec), X)
    assert np.allclose(got['loglikelihood'], a['loglikelihood'], rtol=1e-9, atol=1e-7)
