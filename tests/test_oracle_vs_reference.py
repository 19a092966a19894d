"""Live comparison oracle vs the reference's own code under stubs (build container only; skipped where /root/reference is absent)."""
import numpy as np
import pytest

from oracle import configs, pipeline, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason='/root/reference not present (GPU box)')


@pytest.mark.parametrize('eftlike,solved,mode', [(False, False, None), (True, False, None), (True, True, '.marg'), (True, True, '.best')])
def test_live_reference(eftlike, solved, mode):
    spec = configs.config_kaiser_window(eftlike=eftlike, solved=solved, seed=11)
    if mode == '.best':
        for s in spec['solved']: s['marg'] = False
    like = ref_harness.reference_kaiser_likelihood(spec, eftlike=eftlike, solved_mode=mode)
    if solved: ref_harness.install_jacfwd_stub()
    X = configs.sample_points(spec, 12, seed=5)
    ref = ref_harness.run_vmap(like, spec, X)
    got = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(got['flattheory'] / ref['flattheory'] - 1)) < 1e-12
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-11, atol=2e-9)


def test_reference_vmap_conventions():
    """Shape conventions of the reference's vmap (base.py:131-190): result shaped params.shape + result.shape."""
    ref_harness.setup()
    from desilike.base import vmap
    spec = configs.config_c1()
    like = ref_harness.reference_kaiser_likelihood(spec)
    X = configs.sample_points(spec, 6, seed=1).reshape(2, 3, -1)
    params = {name: X[..., i] for i, name in enumerate(spec['params'])}
    res = vmap(like, backend=None, errors='return')(params)
    assert res[0].shape == (2, 3) and res[1] == {}
    ora = pipeline.evaluate(spec, X.reshape(6, -1))['logposterior'].reshape(2, 3)
    assert np.allclose(res[0], ora, rtol=1e-11, atol=2e-9)


@pytest.mark.parametrize('kwargs,eftlike,mode', [(dict(eftlike=False, solved=False), False, None), (dict(eftlike=True, solved=True), True, '.marg'),
                                                  (dict(eftlike=True, solved=True), True, '.best'), (dict(eftlike=False, solved=False, systematic=True), False, '.marg'),
                                                  (dict(eftlike=False, solved=False, transform='cubic'), False, None)])
def test_spec_from_reference_likelihood(kwargs, eftlike, mode):
    """desilike_b200.spec_from_likelihood: the spec extracted from the reference's OWN initialised likelihood object
    evaluates -- through the oracle here, through the CUDA path on a GPU -- to the reference's own log-posteriors."""
    import desilike_b200 as dl
    spec0 = configs.config_kaiser_window(seed=13, **kwargs)
    if mode == '.best':
        for s in spec0['solved']: s['marg'] = False
    like = ref_harness.reference_kaiser_likelihood(spec0, eftlike=eftlike, solved_mode=mode)
    if spec0['solved']: ref_harness.install_jacfwd_stub()
    X0 = configs.sample_points(spec0, 8, seed=6)
    ref = ref_harness.run_vmap(like, spec0, X0)
    spec = dl.spec_from_likelihood(like)
    assert sorted(spec['params']) == sorted(spec0['params']) and sorted(s['name'] for s in spec['solved']) == sorted(s['name'] for s in spec0['solved'])
    X = np.column_stack([X0[:, spec0['params'].index(name)] for name in spec['params']])
    got = pipeline.evaluate(spec, X, return_theory=True)
    assert np.max(np.abs(got['flattheory'] - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-12
    fin = np.isfinite(ref['logposterior'])
    assert np.array_equal(np.isneginf(ref['logposterior']), np.isneginf(got['logposterior']))
    assert np.allclose(got['logposterior'][fin], ref['logposterior'][fin], rtol=1e-11, atol=2e-9)
    from desilike_b200.lowering import Lowered
    Lowered(spec)   # lowers without error


@pytest.mark.parametrize('which', ['corr', 'power'])
def test_spec_from_reference_bao_likelihood(which):
    """spec_from_likelihood on the reference's damped-BAO xi_ell(s) / P_ell(k) likelihoods with analytically marginalised broadband terms."""
    import desilike_b200 as dl
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import (BAOPowerSpectrumTemplate, DampedBAOWigglesTracerPowerSpectrumMultipoles,
                                                     DampedBAOWigglesTracerCorrelationFunctionMultipoles)
    from desilike.observables.galaxy_clustering import TracerCorrelationFunctionMultipolesObservable, TracerPowerSpectrumMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    rng = np.random.default_rng(9)
    if which == 'corr':
        x = np.arange(20., 152., 4.)
        theory = DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=BAOPowerSpectrumTemplate(z=1.), ells=(0, 2))
        obs = TracerCorrelationFunctionMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, s=[x, x], ells=(0, 2), theory=theory)
        cov, keep = np.diag(np.full(2 * x.size, 1e-8)), ('_0', '_2')
    else:
        x = np.arange(0.02, 0.3, 0.005)
        theory = DampedBAOWigglesTracerPowerSpectrumMultipoles(template=BAOPowerSpectrumTemplate(z=1.), ells=(0, 2))
        obs = TracerPowerSpectrumMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, k=[x, x], ells=(0, 2), theory=theory)
        cov, keep = np.diag(np.full(2 * x.size, 1e4)), ('_-1', '_0', '_1')
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=cov)
    for p in like.all_params:
        if p.basename.startswith('al'):
            if p.basename.endswith(keep): p.update(fixed=False, derived='.marg', prior=dict(dist='norm', loc=0., scale=1e4))
            else: p.update(fixed=True)
    ref_harness.install_jacfwd_stub()
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas']
    X = np.column_stack([rng.uniform(0.97, 1.03, 6), rng.uniform(0.97, 1.03, 6), rng.uniform(1.8, 2.2, 6), rng.uniform(0.9, 1.1, 6), rng.uniform(2., 4., 6)])
    ref = np.array([like(dict(zip(names, xx))) for xx in X])
    spec = dl.spec_from_likelihood(like)
    assert spec['params'] == names and len(spec['solved']) == 2 * len(keep)
    got = pipeline.evaluate(spec, X)
    assert np.allclose(got['logposterior'], ref, rtol=1e-9, atol=1e-7), np.max(np.abs(got['logposterior'] - ref))
    from desilike_b200.lowering import Lowered
    Lowered(spec)


@pytest.mark.parametrize('interp_order', [1, 3])
def test_spec_from_reference_kaiser_correlation_likelihood(interp_order):
    """spec_from_likelihood on the reference's KaiserTracerCorrelationFunctionMultipoles likelihood (full_shape.py:336-365, :553-574), linear and cubic
    (interp_order=3) resampling onto the FFTLog grid: the extracted spec evaluated by the oracle == the reference's own calls, and it lowers."""
    import desilike_b200 as dl
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, KaiserTracerCorrelationFunctionMultipoles
    from desilike.observables.galaxy_clustering import TracerCorrelationFunctionMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    rng = np.random.default_rng(19)
    x = np.arange(20., 152., 4.)
    theory = KaiserTracerCorrelationFunctionMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.), ells=(0, 2), interp_order=interp_order)
    obs = TracerCorrelationFunctionMultipolesObservable(data={'b1': 2.}, s=[x, x], ells=(0, 2), theory=theory)
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=np.diag(np.full(2 * x.size, 1e-8)))
    names = ['qpar', 'qper', 'dm', 'df', 'b1']
    X = np.column_stack([rng.uniform(0.97, 1.03, 6), rng.uniform(0.97, 1.03, 6), rng.uniform(-0.05, 0.05, 6), rng.uniform(0.9, 1.1, 6), rng.uniform(1.8, 2.2, 6)])
    ref = np.array([like(dict(zip(names, xx))) for xx in X])
    spec = dl.spec_from_likelihood(like)
    assert spec['params'] == names and not spec['solved']
    th = spec['observables'][0]['theory']
    assert th['interp_order'] == interp_order and len(th['kin']) == (300 if interp_order == 1 else 100)
    got = pipeline.evaluate(spec, X)
    assert np.allclose(got['logposterior'], ref, rtol=1e-9, atol=1e-7), np.max(np.abs(got['logposterior'] - ref))
    from desilike_b200.lowering import Lowered
    Lowered(spec)


@pytest.mark.parametrize('interp_order', [1, 3])
def test_spec_from_reference_eftlike_kaiser_correlation_likelihood(interp_order):
    """spec_from_likelihood on the reference's EFTLikeKaiserTracerCorrelationFunctionMultipoles likelihood (full_shape.py:577-634, :664-684) with the
    counter terms analytically marginalised: oracle on the extracted spec == the reference's own calls."""
    import desilike_b200 as dl
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, EFTLikeKaiserTracerCorrelationFunctionMultipoles
    from desilike.observables.galaxy_clustering import TracerCorrelationFunctionMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    rng = np.random.default_rng(23)
    x = np.arange(20., 152., 4.)
    theory = EFTLikeKaiserTracerCorrelationFunctionMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.), ells=(0, 2), interp_order=interp_order)
    obs = TracerCorrelationFunctionMultipolesObservable(data={'b1': 2., 'ct0_2': 3., 'ct2_2': -2.}, s=[x, x], ells=(0, 2), theory=theory)
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=np.diag(np.full(2 * x.size, 1e-8)))
    for p in like.all_params:
        if p.basename.startswith('ct'): p.update(fixed=False, derived='.marg', prior=dict(dist='norm', loc=0., scale=50.))
    ref_harness.install_jacfwd_stub()
    names = ['qpar', 'qper', 'dm', 'df', 'b1']
    X = np.column_stack([rng.uniform(0.97, 1.03, 5), rng.uniform(0.97, 1.03, 5), rng.uniform(-0.05, 0.05, 5), rng.uniform(0.9, 1.1, 5), rng.uniform(1.8, 2.2, 5)])
    ref = np.array([like(dict(zip(names, xx))) for xx in X])
    spec = dl.spec_from_likelihood(like)
    assert spec['params'] == names and [s_['name'] for s_ in spec['solved']] == ['ct0_2', 'ct2_2']
    got = pipeline.evaluate(spec, X)
    assert np.allclose(got['logposterior'], ref, rtol=1e-9, atol=1e-6), np.max(np.abs(got['logposterior'] - ref))
    from desilike_b200.lowering import Lowered
    Lowered(spec)


def test_spec_extraction_leaves_the_reference_likelihood_untouched():
    """Extracting a plan must not change the object it reads: ``sub_calculator.all_params`` would build a new pipeline for the sub-calculator
    (reference base.py:949-960) and the likelihood would lose every update made on ``likelihood.all_params`` ('.marg' tags, priors).  The adapter
    reads ``runtime_info.params`` instead: same '.marg' parameters before and after, a second extraction gives the same spec, and the reference
    likelihood still evaluates to the same value."""
    import desilike_b200 as dl
    spec0 = configs.config_kaiser_window(seed=13, eftlike=True, solved=True)
    like = ref_harness.reference_kaiser_likelihood(spec0, eftlike=True, solved_mode='.marg')
    ref_harness.install_jacfwd_stub()
    before = like(b1=1.7)
    tags = {p.name: p.derived for p in like.all_params}
    assert sum(tag == '.marg' for tag in tags.values()) == 7
    first = dl.spec_from_likelihood(like)
    assert {p.name: p.derived for p in like.all_params} == tags
    second = dl.spec_from_likelihood(like)
    assert first['params'] == second['params'] and [s_['name'] for s_ in first['solved']] == [s_['name'] for s_ in second['solved']] and len(first['solved']) == 7
    assert like(b1=1.7) == before


def test_spec_from_reference_likelihood_with_emulated_pt():
    """spec_from_likelihood on a reference likelihood whose ``theory.pt`` is an EmulatedCalculator (reference emulators/__init__.py:394-417): the
    adapter reads ``pt.emulator.__getstate__()`` (the emulator-file dict) into the mirror emulator.  cosmoprimo's emulator classes are absent here,
    so the live reference likelihood (window, data, precision, EFT-like terms with '.marg' parameters: all real reference objects) gets a
    duck-typed ``pt`` carrying the state of a Taylor emulator of the oracle's Kaiser tables.  The extracted spec is a Taylor spec with the
    reference's parameters / solved terms / window; at the expansion point it evaluates to the direct extraction's values, nearby within the
    truncation error, and it lowers (NumPy mirror of the kernels == oracle)."""
    import lowered_numpy
    import desilike_b200 as dl
    from oracle import core
    from desilike_b200.lowering import Lowered
    spec0 = configs.config_kaiser_window(seed=13, eftlike=True, solved=True)
    like = ref_harness.reference_kaiser_likelihood(spec0, eftlike=True, solved_mode='.marg')
    ref_harness.install_jacfwd_stub()
    like(b1=1.7)
    direct = dl.spec_from_likelihood(like)
    th = direct['observables'][0]['theory']
    pnames, center, delta = ['qpar', 'qper', 'dm', 'df'], np.array([1., 1., 0., 1.]), np.array([0.01, 0.01, 0.01, 0.02])
    ptheory = dict(th, params=pnames)

    def func(P):
        tabs = pipeline.theory_tables(ptheory, P, pnames)
        return np.concatenate([tabs[n] for n in ('pk_dd', 'pk_dt', 'pk_tt')], axis=-1)
    powers, derivs = core.taylor_fit_polynomial(func, center, 4, 2, delta)
    nell, nk = len(th['ells']), len(th['k'])
    tables = {n: derivs[:, i * nell * nk:(i + 1) * nell * nk].reshape(-1, nell, nk) for i, n in enumerate(('pk_dd', 'pk_dt', 'pk_tt'))}
    state = dl.Emulator.from_taylor(pnames, center, powers, tables, fixed=dict(k=th['k'], ells=tuple(th['ells']), z=1.)).state

    class FakeEmulator(object):
        def __getstate__(self): return state

    class FakePT(object):
        emulator = FakeEmulator()

    def proxy(obj, **override):   # (assigning to a reference calculator re-initialises its pipeline: read-only proxies instead)
        p_ = type(type(obj).__name__, (object,), {'__getattr__': lambda self, name: getattr(obj, name)})()
        p_.__dict__.update(override)
        return p_

    obs = like.observables[0]
    spec = dl.spec_from_likelihood(proxy(like, observables=[proxy(obs, wmatrix=proxy(obs.wmatrix, theory=proxy(obs.wmatrix.theory, pt=FakePT())))]))
    assert spec['observables'][0]['theory']['kind'] == 'taylor'
    assert spec['params'] == direct['params'] and [s_['name'] for s_ in spec['solved']] == [s_['name'] for s_ in direct['solved']]
    names = spec['params']
    x0 = np.array([dict(zip(pnames, center)).get(n, 1.7 if n == 'b1' else 0.) for n in names])[None, :]
    a, b = pipeline.evaluate(spec, x0), pipeline.evaluate(direct, x0)
    assert np.allclose(a['logposterior'], b['logposterior'], rtol=1e-9, atol=1e-7) and np.allclose(a['x_solved'], b['x_solved'], rtol=1e-7, atol=1e-7)
    rng = np.random.default_rng(2)
    scale = np.array([0.3 * dict(zip(pnames, delta)).get(n, 0.3) for n in names])
    X = x0 + rng.uniform(-1., 1., (6, len(names))) * scale
    a, b = pipeline.evaluate(spec, X, return_theory=True), pipeline.evaluate(direct, X, return_theory=True)
    assert np.max(np.abs(a['flattheory'] - b['flattheory'])) < 1e-4 * np.max(np.abs(b['flattheory']))
    got = lowered_numpy.evaluate(Lowered(spec), X)
    assert np.allclose(got['logposterior'], a['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.allclose(got['x_solved'], a['x_solved'], rtol=1e-7, atol=1e-8 * np.max(np.abs(a['x_solved'])))


@pytest.mark.parametrize('which', ['power', 'corr'])
def test_spec_from_reference_bao_likelihood_with_emulated_pt(which):
    """Same for the damped-BAO likelihoods: ``theory.pt`` = an EmulatedCalculator with a 'power' engine -- on the observable's k for P_ell(k), on kin
    behind get_corr for xi_ell(s) (reference bao.py:820-833) -- duck-typed, carrying the state of a Taylor emulator of the oracle's damped-BAO
    multipoles; real reference window / data / precision / marginalised broadband terms."""
    import lowered_numpy
    import desilike_b200 as dl
    from oracle import core
    from desilike_b200.lowering import Lowered
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import (BAOPowerSpectrumTemplate, DampedBAOWigglesTracerPowerSpectrumMultipoles,
                                                     DampedBAOWigglesTracerCorrelationFunctionMultipoles)
    from desilike.observables.galaxy_clustering import TracerPowerSpectrumMultipolesObservable, TracerCorrelationFunctionMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    if which == 'power':
        x = np.arange(0.02, 0.3, 0.005)
        theory = DampedBAOWigglesTracerPowerSpectrumMultipoles(template=BAOPowerSpectrumTemplate(z=1.), ells=(0, 2))
        obs = TracerPowerSpectrumMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, k=[x, x], ells=(0, 2), theory=theory)
        cov, keep = np.diag(np.full(2 * x.size, 1e4)), ('_-1', '_0', '_1')
    else:
        x = np.arange(20., 152., 4.)
        theory = DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=BAOPowerSpectrumTemplate(z=1.), ells=(0, 2))
        obs = TracerCorrelationFunctionMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, s=[x, x], ells=(0, 2), theory=theory)
        cov, keep = np.diag(np.full(2 * x.size, 1e-8)), ('_0', '_2')
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=cov)
    for p in like.all_params:
        if p.basename.startswith('al'):
            if p.basename.endswith(keep): p.update(fixed=False, derived='.marg', prior=dict(dist='norm', loc=0., scale=1e4))
            else: p.update(fixed=True)
    ref_harness.install_jacfwd_stub()
    like(b1=2.)
    direct = dl.spec_from_likelihood(like)
    th = direct['observables'][0]['theory']
    pnames, center, delta = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas'], np.array([1., 1., 2., 1., 3.]), np.array([0.01, 0.01, 0.05, 0.02, 0.2])
    # the emulated calculator: P_ell on the observable's k, or on kin (the P(k) theory behind get_corr)
    kem = x if which == 'power' else np.asarray(th['kin'], dtype='f8')
    ptheory = dict({key: value for key, value in th.items() if key not in ('s', 'kin', 'interp_order', 'fftlog')}, kind='bao_pk', k=kem, params=pnames)
    powers, derivs = core.taylor_fit_polynomial(lambda P: pipeline.theory_tables(ptheory, P, pnames)['power'], center, 5, 2, delta)
    state = dl.Emulator.from_taylor(pnames, center, powers, {'power': derivs.reshape(-1, 2, kem.size)}, fixed=dict(k=kem, ells=(0, 2), z=1.),
                                    calculator_class=('desilike.theories.galaxy_clustering.bao', 'DampedBAOWigglesPowerSpectrumMultipoles')).state

    class FakeEmulator(object):
        def __getstate__(self): return state

    class FakePT(object):
        emulator = FakeEmulator()

    def proxy(obj, **override):
        p_ = type(type(obj).__name__, (object,), {'__getattr__': lambda self, name: getattr(obj, name)})()
        p_.__dict__.update(override)
        return p_

    o = like.observables[0]
    spec = dl.spec_from_likelihood(proxy(like, observables=[proxy(o, wmatrix=proxy(o.wmatrix, theory=proxy(o.wmatrix.theory, pt=FakePT())))]))
    assert spec['observables'][0]['theory']['kind'] == 'taylor' and spec['params'] == direct['params'] == pnames
    assert [s_['name'] for s_ in spec['solved']] == [s_['name'] for s_ in direct['solved']] and len(spec['solved']) == 2 * len(keep)
    a, b = pipeline.evaluate(spec, center[None, :]), pipeline.evaluate(direct, center[None, :])
    assert np.allclose(a['logposterior'], b['logposterior'], rtol=1e-9, atol=1e-7)
    rng = np.random.default_rng(4)
    X = center + rng.uniform(-1., 1., (6, 5)) * 0.3 * delta
    a, b = pipeline.evaluate(spec, X, return_theory=True), pipeline.evaluate(direct, X, return_theory=True)
    assert np.max(np.abs(a['flattheory'] - b['flattheory'])) < 1e-4 * np.max(np.abs(b['flattheory']))
    got = lowered_numpy.evaluate(Lowered(spec), X)
    assert np.allclose(got['logposterior'], a['logposterior'], rtol=1e-9, atol=1e-7)


def test_vmap_accepts_reference_likelihood_objects():
    """desilike_b200.vmap(reference_likelihood): the adapter extracts the spec and maps parameter dicts to plan columns (no GPU here: the plan
    itself is not built; the point matrix and the spec are checked)."""
    from desilike_b200.base import ReferenceLikelihoodAdapter
    import desilike_b200 as dl
    spec0 = configs.config_kaiser_window(eftlike=True, solved=True, seed=13)
    like = ref_harness.reference_kaiser_likelihood(spec0, eftlike=True, solved_mode='.marg')
    ref_harness.install_jacfwd_stub()
    like(b1=1.7)
    adapter = ReferenceLikelihoodAdapter(like)
    X = adapter._points({'b1': np.array([1.5, 2.]), 'qpar': np.array([1.01, 0.99])})
    names = adapter.spec['params']
    assert X.shape == (2, len(names)) and np.array_equal(X[:, names.index('b1')], [1.5, 2.]) and np.all(X[:, names.index('qper')] == 1.)
    with pytest.raises(dl.PipelineError):
        adapter._points({'nope': np.zeros(2)})
    wrapper = dl.vmap(like, backend='cuda')   # wrapping needs no GPU; evaluation does
    assert callable(wrapper)


def test_spec_from_reference_sum_likelihood():
    """SumLikelihood of two reference likelihoods sharing their parameters -> one spec with one precision block per member."""
    import desilike_b200 as dl
    spec0 = configs.config_kaiser_window(eftlike=False, solved=False, seed=13)
    like1 = ref_harness.reference_kaiser_likelihood(spec0)
    like2 = ref_harness.reference_kaiser_likelihood(configs.config_kaiser_window(eftlike=False, solved=False, seed=14))
    both = like1 + like2
    X0 = configs.sample_points(spec0, 5, seed=2)
    ref = np.array([both(dict(zip(spec0['params'], x))) for x in X0])
    spec = dl.spec_from_likelihood(both)
    assert len(spec['observables']) == 2 and len(spec['precision']) == 2
    X = np.column_stack([X0[:, spec0['params'].index(name)] for name in spec['params']])
    got = pipeline.evaluate(spec, X)
    fin = np.isfinite(ref)
    assert np.allclose(got['logposterior'][fin], ref[fin], rtol=1e-11, atol=2e-9)


def test_spec_from_reference_likelihood_with_k_mask():
    """Reference observable defined through ``klim`` with different k ranges per multipole (window ``kmask``, reference window.py:296-303, :466-467)."""
    import desilike_b200 as dl
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, KaiserTracerPowerSpectrumMultipoles
    from desilike.observables.galaxy_clustering import TracerPowerSpectrumMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    theory = KaiserTracerPowerSpectrumMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.))
    klim = {0: [0.02, 0.2, 0.01], 2: [0.04, 0.16, 0.01], 4: [0.06, 0.1, 0.01]}
    obs = TracerPowerSpectrumMultipolesObservable(data={'b1': 2.}, klim=klim, theory=theory)
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=np.diag(np.full(18 + 12 + 4, 1e5)))
    rng = np.random.default_rng(5)
    names = ['qpar', 'qper', 'dm', 'df', 'b1', 'sn0']
    X = np.column_stack([rng.uniform(0.97, 1.03, 5), rng.uniform(0.97, 1.03, 5), rng.uniform(-0.05, 0.05, 5), rng.uniform(0.9, 1.1, 5), rng.uniform(1.8, 2.2, 5), rng.normal(0., 0.1, 5)])
    ref = np.array([like(dict(zip(names, x))) for x in X])
    assert obs.wmatrix.kmask is not None
    spec = dl.spec_from_likelihood(like)
    assert spec['observables'][0]['wmatrix'].shape[0] == len(obs.flatdata) < spec['observables'][0]['nin']
    got = pipeline.evaluate(spec, np.column_stack([X[:, names.index(n)] for n in spec['params']]))
    assert np.allclose(got['logposterior'], ref, rtol=1e-10, atol=1e-8)


def test_spec_from_reference_xi_likelihood_with_fiber_collisions():
    """Reference xi observable with its top-hat fiber-collision window (reference window.py:686-702, :1192-1249): the adapter reads matrix_diag / offset,
    and the mirror observable built with ``fiber_collisions=dict(fs, Dfc)`` folds to the same matrix and offset."""
    import desilike_b200 as dl
    from oracle import synthetic
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import BAOPowerSpectrumTemplate, DampedBAOWigglesTracerCorrelationFunctionMultipoles
    from desilike.observables.galaxy_clustering import TracerCorrelationFunctionMultipolesObservable, TopHatFiberCollisionsCorrelationFunctionMultipoles
    from desilike.likelihoods import ObservablesGaussianLikelihood
    s = np.arange(20., 152., 4.)
    theory = DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=BAOPowerSpectrumTemplate(z=1.), ells=(0, 2))
    fc = TopHatFiberCollisionsCorrelationFunctionMultipoles(fs=0.5, Dfc=12.)
    obs = TracerCorrelationFunctionMultipolesObservable(data={'b1': 2., 'sigmas': 3.}, s=[s, s], ells=(0, 2), theory=theory, fiber_collisions=fc)
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=np.diag(np.full(66, 1e-8)))
    for p in like.all_params:
        if p.basename.startswith('al'): p.update(fixed=True)
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas']
    rng = np.random.default_rng(12)
    X = np.column_stack([rng.uniform(0.97, 1.03, 4), rng.uniform(0.97, 1.03, 4), rng.uniform(1.8, 2.2, 4), rng.uniform(0.9, 1.1, 4), rng.uniform(2., 4., 4)])
    ref = np.array([like(dict(zip(names, x))) for x in X])
    assert obs.wmatrix.matrix_diag is not None
    spec = dl.spec_from_likelihood(like)
    got = pipeline.evaluate(spec, X)
    assert np.allclose(got['logposterior'], ref, rtol=1e-9, atol=1e-7)
    template = dl.BAOPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID, rs_drag=synthetic.RS_DRAG)
    mirror = dl.TracerCorrelationFunctionMultipolesObservable(data=np.zeros(66), s=[s, s], ells=(0, 2), fiber_collisions=dict(fs=0.5, Dfc=12.),
                                                              theory=dl.DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=template, ells=(0, 2)))
    ospec = mirror.observable_spec()
    assert np.allclose(ospec['wmatrix'], spec['observables'][0]['wmatrix'], rtol=0., atol=1e-14) and np.allclose(ospec['offset'], spec['observables'][0]['offset'], rtol=0., atol=1e-14)


def test_reference_power_observable_with_fiber_collisions():
    """Reference P(k) observable with its top-hat fiber-collision window (reference window.py:428-438, :967-1049; ``utils.weights_trapz`` shimmed with its
    jax semantics, see tests/golden/make_golden.py:fiber_collisions): adapter spec == reference log-posterior; mirror fold == the reference's matrix / offset."""
    import desilike_b200 as dl
    from oracle import synthetic
    ref_harness.setup()
    from desilike import utils as ref_utils
    ref_utils.weights_trapz = lambda x: np.concatenate([[x[1] - x[0]], x[2:] - x[:-2], [x[-1] - x[-2]]]) / 2.
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, KaiserTracerPowerSpectrumMultipoles
    from desilike.observables.galaxy_clustering import TracerPowerSpectrumMultipolesObservable, TopHatFiberCollisionsPowerSpectrumMultipoles
    from desilike.likelihoods import ObservablesGaussianLikelihood
    k = np.arange(0.02, 0.2, 0.005)
    theory = KaiserTracerPowerSpectrumMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.))
    fc = TopHatFiberCollisionsPowerSpectrumMultipoles(fs=0.5, Dfc=0.8)
    obs = TracerPowerSpectrumMultipolesObservable(data={'b1': 2.}, k=[k] * 3, ells=(0, 2, 4), theory=theory, fiber_collisions=fc)
    like = ObservablesGaussianLikelihood(observables=[obs], covariance=np.diag(np.full(3 * k.size, 1e5)))
    names = ['qpar', 'qper', 'dm', 'df', 'b1', 'sn0']
    rng = np.random.default_rng(15)
    X = np.column_stack([rng.uniform(0.97, 1.03, 4), rng.uniform(0.97, 1.03, 4), rng.uniform(-0.05, 0.05, 4), rng.uniform(0.9, 1.1, 4), rng.uniform(1.8, 2.2, 4), rng.normal(0., 0.1, 4)])
    ref = np.array([like(dict(zip(names, x))) for x in X])
    spec = dl.spec_from_likelihood(like)
    got = pipeline.evaluate(spec, np.column_stack([X[:, names.index(n)] for n in spec['params']]))
    assert np.allclose(got['logposterior'], ref, rtol=1e-10, atol=1e-8)
    template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=synthetic.pk_fid, pknow_dd_fid=synthetic.pknow_fid, f_fid=synthetic.F_FID)
    mirror = dl.TracerPowerSpectrumMultipolesObservable(data=np.zeros(3 * k.size), k=[k] * 3, ells=(0, 2, 4), fiber_collisions=dict(fs=0.5, Dfc=0.8),
                                                        theory=dl.KaiserTracerPowerSpectrumMultipoles(template=template))
    assert np.allclose(mirror.matrix_full, np.asarray(obs.wmatrix.matrix_full), rtol=1e-12, atol=1e-14)
    assert np.allclose(mirror.offset, np.asarray(obs.wmatrix.offset), rtol=1e-10, atol=1e-9)


def test_qmc_design_matches_reference_sequence():
    """``Emulator.qmc_points`` (training design of MLP emulators) == the reference's in-tree ``RQuasiRandomSequence`` + ``qmc.scale``
    (desilike/samplers/qmc.py:13-36, :122-126), including a resumed sequence (offset)."""
    ref_harness.setup()
    from scipy.stats import qmc
    from desilike.samplers.qmc import RQuasiRandomSequence
    import desilike_b200 as dl
    lower, upper = np.array([0.9, 0.9, -0.1, 0.5]), np.array([1.1, 1.2, 0.1, 1.5])
    eng = RQuasiRandomSequence(d=4, seed=0.5)
    eng.reset()
    ref = qmc.scale(eng.random(n=257), lower, upper)
    got = dl.Emulator.qmc_points(lower, upper, 257, engine='rqrs', seed=0.5)
    assert np.allclose(got, ref, rtol=0, atol=1e-15)
    eng.reset(); eng.fast_forward(100)
    ref = qmc.scale(eng.random(n=50), lower, upper)
    assert np.allclose(dl.Emulator.qmc_points(lower, upper, 50, offset=100), ref, rtol=0, atol=1e-15)
