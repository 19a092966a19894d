"""Generate tests/golden/*.npz: outputs of the reference's OWN NumPy code path (under stub third-party modules) on
seeded inputs.  Runs only in the build container (needs /root/reference); the fixtures are committed and travel.

    python tests/golden/make_golden.py

Each fixture stores the config name + kwargs (the spec is rebuilt from its seed by oracle/configs.py), the points X and the
reference's logposterior / loglikelihood / logprior / flattheory.  See oracle/ref_harness.py for how the reference is run.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import configs, ref_harness  # noqa: E402

CASES = {
    'c1_kaiser_shapefit': dict(config='config_c1', kwargs={}, eftlike=False, mode=None, n=64, seed=101),
    'kaiser_window': dict(config='config_kaiser_window', kwargs=dict(eftlike=False, solved=False), eftlike=False, mode=None, n=48, seed=102),
    'eftlike_sampled': dict(config='config_kaiser_window', kwargs=dict(eftlike=True, solved=False), eftlike=True, mode=None, n=48, seed=103),
    'eftlike_marg': dict(config='config_kaiser_window', kwargs=dict(eftlike=True, solved=True), eftlike=True, mode='.marg', n=48, seed=104),
    'eftlike_best': dict(config='config_kaiser_window', kwargs=dict(eftlike=True, solved=True), eftlike=True, mode='.best', n=48, seed=105),
    'eftlike_prec': dict(config='config_kaiser_window', kwargs=dict(eftlike=True, solved=True), eftlike=True, mode='.prec', n=48, seed=109),
    'kaiser_window_syst': dict(config='config_kaiser_window', kwargs=dict(eftlike=False, solved=False, systematic=True), eftlike=False, mode='.marg', n=32, seed=111),
    'kaiser_window_cubic': dict(config='config_kaiser_window', kwargs=dict(eftlike=False, solved=False, transform='cubic'), eftlike=False, mode=None, n=32, seed=106),
    # transform='cubic' together with analytically marginalised terms: flatdiff is not linear in them, the reference linearises at the defaults
    # (jacfwd through the transform, likelihoods/base.py:152-158) -- the jacfwd shim then uses an extrapolated small-step difference
    'eftlike_cubic_marg': dict(config='config_kaiser_window', kwargs=dict(eftlike=True, solved=True, transform='cubic'), eftlike=True, mode='.marg', n=32, seed=112, jac_step=3e-3),
}


def build_spec(case):
    spec = getattr(configs, case['config'])(**case['kwargs'])
    if case['mode'] == '.best':
        for s in spec['solved']: s['marg'] = False
    if case['mode'] == '.prec':
        for s in spec['solved']: s['mode'] = 'prec'
    return spec


def main(only=None):
    for name, case in CASES.items():
        if only and name not in only: continue
        spec = build_spec(case)
        like = ref_harness.reference_kaiser_likelihood(spec, eftlike=case['eftlike'], solved_mode=case['mode'])
        if spec['solved']: ref_harness.install_jacfwd_stub(step=case.get('jac_step', 1.))
        X = configs.sample_points(spec, case['n'], seed=case['seed'])
        if not spec['solved']:
            X[1, 0] = 1.3   # out of prior -> -inf
        ref = ref_harness.run_vmap(like, spec, X)
        if case['mode'] == '.prec':   # the Jacobian of the pre-marginalisation is taken at the reference's default parameter values
            case = dict(case, fiducial={p.name: float(p.value) for p in like.all_params if p.name in spec['params']})
        np.savez_compressed(os.path.join(HERE, name + '.npz'), case=json.dumps(case), X=X, **ref)
        print(name, X.shape, ref['logposterior'][:3])


def bao_theory():
    """Theory-level fixture: the reference's damped-BAO P_ell(k) class and its P -> xi class (get_corr wiring; the FFTLog
    call itself is routed to oracle/fftlog.py by the cosmoprimo stub) on seeded points."""
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import (BAOPowerSpectrumTemplate, DampedBAOWigglesTracerPowerSpectrumMultipoles,
                                                     DampedBAOWigglesTracerCorrelationFunctionMultipoles)
    k, s, ells = np.arange(0.02, 0.3, 0.005), np.arange(20., 152., 4.), (0, 2)
    pk = DampedBAOWigglesTracerPowerSpectrumMultipoles(template=BAOPowerSpectrumTemplate(z=1.), k=k, ells=ells)
    xi = DampedBAOWigglesTracerCorrelationFunctionMultipoles(template=BAOPowerSpectrumTemplate(z=1.), s=s, ells=ells)
    names = ['qpar', 'qper', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper']
    spec = dict(params=names)
    X = configs.sample_points(spec, 12, seed=107)
    power = np.array([pk(**dict(zip(names, x))) for x in X])
    corr = np.array([xi(**dict(zip(names, x))) for x in X])
    np.savez_compressed(os.path.join(HERE, 'bao_theory.npz'), case=json.dumps(dict(kind='bao_theory')), X=X, k=k, s=s, power=power, corr=corr,
                        k_template_pk=np.array(pk.template.k), k_template_xi=np.array(xi.template.k))
    print('bao_theory', power.shape, corr.shape)
    # broadband terms (reference bao.py:467-476, :527-534 power law in k / kp; :849-853, :867, :905 in s / sp)
    rng = np.random.default_rng(114)
    bb_pk, bb_xi = ['al0_-1', 'al0_0', 'al2_1', 'al2_-3'], ['al0_0', 'al0_2', 'al2_-2', 'al2_1']
    A_pk, A_xi = rng.uniform(-50., 50., (X.shape[0], len(bb_pk))), rng.uniform(-1e-3, 1e-3, (X.shape[0], len(bb_xi)))
    power_bb = np.array([pk(**dict(zip(names, x)), **dict(zip(bb_pk, a))) for x, a in zip(X, A_pk)])
    corr_bb = np.array([xi(**dict(zip(names, x)), **dict(zip(bb_xi, a))) for x, a in zip(X, A_xi)])
    np.savez_compressed(os.path.join(HERE, 'bao_broadband.npz'), case=json.dumps(dict(kind='bao_broadband', bb_pk=bb_pk, bb_xi=bb_xi)), X=X, k=k, s=s,
                        A_pk=A_pk, A_xi=A_xi, power=power_bb, corr=corr_bb, kp=float(pk.kp), sp=float(xi.sp))
    print('bao_broadband', power_bb.shape, corr_bb.shape, float(pk.kp), float(xi.sp))


def lpt_combine():
    """Function-level fixture: the reference's ``tablevel_combine_bias_terms_poles`` (full_shape.py:1182-1186) and its ``interp1d``
    (jax.py:211-265, NumPy fallback) as chained in ``LPTVelocileptorsTracerPowerSpectrumMultipoles.calculate`` (:1310-1312), on a
    seeded random pktable (velocileptors itself is a third-party code, absent here)."""
    ref_harness.setup()
    from desilike.theories.galaxy_clustering.full_shape import tablevel_combine_bias_terms_poles
    from desilike.jax import interp1d
    rng = np.random.default_rng(110)
    nell = 3
    kpt = np.concatenate([[0.0005], np.geomspace(0.0015, 0.025, 20, endpoint=True), np.arange(0.03, 0.5 + 0.0075, 0.005)])   # full_shape.py:1283
    k = np.arange(0.02, 0.2, 0.005)
    pktable = rng.standard_normal((nell, kpt.size, 19)) * (1e4 * np.exp(-kpt / 0.2))[None, :, None]
    pars = np.column_stack([rng.uniform(0.5, 2., 8), rng.uniform(-2., 2., 8), rng.uniform(-2., 2., 8), rng.uniform(-1., 1., 8)] + [rng.uniform(-20., 20., 8) for i in range(4)]
                           + [rng.uniform(-1., 1., 8) for i in range(3)])
    nd = 1e-4
    table = np.array([np.asarray(tablevel_combine_bias_terms_poles(pktable, list(p), nd=nd)) for p in pars])
    power = np.array([np.asarray(interp1d(k, kpt, t.T)).T for t in table])
    np.savez_compressed(os.path.join(HERE, 'lpt_combine.npz'), case=json.dumps(dict(kind='lpt_combine')), pktable=pktable, kpt=kpt, k=k, pars=pars, nd=nd,
                        table=table, power=power)
    print('lpt_combine', table.shape, power.shape)


def kaiser_corr():
    """Theory-level fixture: the reference's KaiserTracerCorrelationFunctionMultipoles (full_shape.py:336-365, :553-574: Hankel transforms of the Kaiser
    tracer P_ell on kin through get_corr, base.py:46-136; FFTLog routed to oracle/fftlog.py by the cosmoprimo stub) with the linear and the cubic
    (interp_order=3) resampling onto the FFTLog grid, on seeded points."""
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, KaiserTracerCorrelationFunctionMultipoles
    s, ells = np.arange(20., 152., 4.), (0, 2)
    names = ['qpar', 'qper', 'dm', 'df', 'b1']
    rng = np.random.default_rng(117)
    X = np.stack([rng.uniform(0.9, 1.1, 10), rng.uniform(0.9, 1.1, 10), rng.uniform(-0.1, 0.1, 10), rng.uniform(0.8, 1.2, 10), rng.uniform(1., 3., 10)], axis=-1)
    out = {}
    for order in (1, 3):
        xi = KaiserTracerCorrelationFunctionMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.), s=s, ells=ells, interp_order=order)
        out['corr{:d}'.format(order)] = np.array([xi(**dict(zip(names, x))) for x in X])
        out['kin{:d}'.format(order)] = np.array(xi.kin)
        out['k_template'] = np.array(xi.template.k)
        out['a'], out['kp'] = float(xi.template.a), float(xi.template.kp)
    # EFT-like class (full_shape.py:577-634, :664-684): counter terms ct{ell}_2 (k / kp)^2 pk11_{ell=0} on kin, transformed with the rest
    from desilike.theories.galaxy_clustering import EFTLikeKaiserTracerCorrelationFunctionMultipoles
    ct = rng.uniform(-20., 20., (X.shape[0], 2))
    for order in (1, 3):
        xi = EFTLikeKaiserTracerCorrelationFunctionMultipoles(template=ShapeFitPowerSpectrumTemplate(z=1.), s=s, ells=ells, interp_order=order)
        out['corr_eft{:d}'.format(order)] = np.array([xi(**dict(zip(names + ['ct0_2', 'ct2_2'], np.append(x, c)))) for x, c in zip(X, ct)])
    out['ct'] = ct
    np.savez_compressed(os.path.join(HERE, 'kaiser_corr.npz'), case=json.dumps(dict(kind='kaiser_corr', names=names)), X=X, s=s, **out)
    print('kaiser_corr', out['corr1'].shape, float(np.max(np.abs(out['corr3'] - out['corr1'])) / np.max(np.abs(out['corr1']))))


def fiber_collisions():
    """Class-level fixture: kernels of the reference's TopHatFiberCollisionsCorrelationFunctionMultipoles (window.py:1192-1249), instantiated
    with a stub theory that only carries (s, ells)."""
    ref_harness.setup()
    from desilike.observables.galaxy_clustering.window import TopHatFiberCollisionsCorrelationFunctionMultipoles as Cls

    class Theory(object):
        s, ells = np.arange(2., 80., 3.), (0, 2, 4)

        class init(object):
            update = staticmethod(lambda **kw: None)

        class runtime_info(object):
            initialize = staticmethod(lambda: None)

    out = {}
    for name, kw in {'a': dict(fs=0.6, Dfc=3., mu_range_cut=False), 'b': dict(fs=0.35, Dfc=10., mu_range_cut=True)}.items():
        fc = object.__new__(Cls)
        Cls.initialize(fc, ells=(0, 2), theory=Theory(), **kw)
        out['correlated_' + name], out['uncorrelated_' + name] = np.array(fc.kernel_correlated), np.array(fc.kernel_uncorrelated)
    from desilike.observables.galaxy_clustering.window import TopHatFiberCollisionsPowerSpectrumMultipoles as ClsP

    class TheoryP(object):
        k, ells = np.linspace(0.005, 0.4, 60), (0, 2, 4)

        class runtime_info(object):
            initialize = staticmethod(lambda: None)

    # the reference's utils.weights_trapz (utils.py:614-622) relies on jax.numpy.insert clipping the out-of-range index len(x) - 1 (NumPy raises):
    # shim it with the jax semantics (the last weight is appended) so that the reference class runs unmodified in the NumPy fallback
    from desilike import utils as ref_utils
    ref_utils.weights_trapz = lambda x: np.concatenate([[x[1] - x[0]], x[2:] - x[:-2], [x[-1] - x[-2]]]) / 2.
    fcp = object.__new__(ClsP)
    ClsP.initialize(fcp, k=TheoryP.k, ells=(0, 2, 4), theory=TheoryP(), fs=0.5, Dfc=0.8)
    out['p_correlated'], out['p_uncorrelated'], out['p_k'] = np.array(fcp.kernel_correlated), np.array(fcp.kernel_uncorrelated), TheoryP.k
    np.savez_compressed(os.path.join(HERE, 'fiber_collisions.npz'), case=json.dumps(dict(kind='fiber_collisions', a=dict(fs=0.6, Dfc=3., mu_range_cut=False), b=dict(fs=0.35, Dfc=10., mu_range_cut=True))),
                        s=Theory.s, ells=np.array([0, 2]), ellsin=np.array(Theory.ells), **out)
    print('fiber_collisions', out['correlated_a'].shape)


def covariance_corrections():
    """Likelihood-level fixture: the reference's ObservablesGaussianLikelihood built from a COVARIANCE with ``scale_covariance`` and the
    Hartlap 2007 / Percival 2014 corrections (likelihoods/base.py:605-656) on the EFT-like marginalised configuration."""
    ref_harness.setup()
    from desilike.theories.galaxy_clustering import ShapeFitPowerSpectrumTemplate, EFTLikeKaiserTracerPowerSpectrumMultipoles
    from desilike.observables.galaxy_clustering import TracerPowerSpectrumMultipolesObservable
    from desilike.likelihoods import ObservablesGaussianLikelihood
    spec = configs.config_kaiser_window(eftlike=True, solved=True)
    obs = spec['observables'][0]
    th = obs['theory']
    template = ShapeFitPowerSpectrumTemplate(z=1., a=th['template']['a'], kp=th['template']['kp'], apmode=th['template']['apmode'])
    theory = EFTLikeKaiserTracerPowerSpectrumMultipoles(template=template, mu=th['nmu'])
    nell = len(th['ells'])
    observable = TracerPowerSpectrumMultipolesObservable(data=np.asarray(obs['data']), theory=theory, wmatrix=np.asarray(obs['wmatrix_raw']), kin=th['k'], ellsin=th['ells'],
                                                         k=[obs['kout']] * nell, ells=th['ells'], shotnoise=obs.get('shotnoise'))
    cov = np.linalg.inv(spec['precision'][0])
    nobs, scale = 400, 1.3
    like = ObservablesGaussianLikelihood(observables=[observable], covariance=cov, scale_covariance=scale, correct_covariance={'nobs': nobs, 'correction': 'hartlap-percival2014'})
    for param in like.all_params:
        name = param.name
        if name in spec['params']: param.update(fixed=False, prior=dict(spec['priors'][name]) if spec['priors'].get(name) else None)
        elif any(s_['name'] == name for s_ in spec['solved']):
            s_ = [s_ for s_ in spec['solved'] if s_['name'] == name][0]
            param.update(fixed=False, value=s_['value'], prior=dict(s_['prior']), derived='.marg')
        elif not param.derived: param.update(fixed=True)
    ref_harness.install_jacfwd_stub()
    X = configs.sample_points(spec, 16, seed=113)
    ref = ref_harness.run_vmap(like, spec, X)
    np.savez_compressed(os.path.join(HERE, 'covariance_corrections.npz'), case=json.dumps(dict(kind='covariance_corrections', nobs=nobs, scale_covariance=scale)), X=X,
                        covariance=cov, hartlap2007_factor=like.hartlap2007_factor, percival2014_factor=like.percival2014_factor, precision=np.array(like.precision), **ref)
    print('covariance_corrections', like.hartlap2007_factor, like.percival2014_factor)


def xi_window():
    """Method-level fixture: the reference's ``WindowedCorrelationFunctionMultipoles._apply`` (window.py:717-728) called on an instance carrying
    only the constants it reads (matrix_diag, matrix_full, offset, smask), for the two combinations the mirror class folds."""
    ref_harness.setup()
    from desilike.observables.galaxy_clustering.window import WindowedCorrelationFunctionMultipoles
    rng = np.random.default_rng(112)
    sin = np.arange(20., 152., 4.)
    ns = sin.size
    s0, s1 = sin[:-2], sin[3:]
    smask = np.concatenate([np.searchsorted(sin, s0), ns + np.searchsorted(sin, s1)])
    nout = smask.size
    matrix_diag = np.array([[np.ones(ns) - 0.02, 0.01 * np.ones(ns)], [0.03 * sin / 100., np.ones(ns) * 0.97]])
    theory = rng.standard_normal((4, 2, ns))
    out = {}
    for name, (mf, off, mask) in {'diag': (None, 1e-4 * rng.standard_normal(2 * ns), smask), 'full': (rng.standard_normal((nout, 2 * ns)), 1e-4 * rng.standard_normal(nout), None)}.items():
        w = object.__new__(WindowedCorrelationFunctionMultipoles)
        for key, value in dict(matrix_diag=matrix_diag, matrix_full=mf, offset=off, smask=mask).items(): object.__setattr__(w, key, value)
        out['out_' + name] = np.array([np.asarray(WindowedCorrelationFunctionMultipoles._apply(w, th)) for th in theory])
        out['offset_' + name] = off
        if mf is not None: out['matrix_full'] = mf
    np.savez_compressed(os.path.join(HERE, 'xi_window.npz'), case=json.dumps(dict(kind='xi_window')), sin=sin, s0=s0, s1=s1, matrix_diag=matrix_diag, theory=theory, **out)
    print('xi_window', out['out_diag'].shape, out['out_full'].shape)


if __name__ == '__main__':
    if len(sys.argv) > 1:   # python tests/golden/make_golden.py <case> [<case> ...]: (re)generate only those likelihood fixtures
        main(only=sys.argv[1:])
        sys.exit(0)
    fiber_collisions()
    covariance_corrections()
    xi_window()
    lpt_combine()
    bao_theory()
    kaiser_corr()
    main()
