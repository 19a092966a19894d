"""The oracle against the committed golden fixtures (outputs of the reference's own NumPy path, tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import configs, pipeline

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(f for f in glob.glob(os.path.join(HERE, 'golden', '*.npz')) if 'bao_theory' not in f and 'bao_broadband' not in f and 'lpt_combine' not in f and 'xi_window' not in f and 'covariance_corrections' not in f and 'fiber_collisions' not in f and 'taylor_state' not in f and 'kaiser_corr' not in f)


def load_case(fn):
    data = np.load(fn)
    case = json.loads(str(data['case']))
    spec = getattr(configs, case['config'])(**case['kwargs'])
    if case['mode'] == '.best':
        for s in spec['solved']: s['marg'] = False
    if case['mode'] == '.prec':
        for s in spec['solved']: s['mode'] = 'prec'
        spec['fiducial'] = case['fiducial']
    return case, spec, data


def test_fixtures_present():
    assert len(FIXTURES) >= 7


@pytest.mark.parametrize('fn', FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_oracle_matches_reference_fixture(fn):
    case, spec, data = load_case(fn)
    X = data['X']
    got = pipeline.evaluate(spec, X, return_theory=True)
    scale = np.max(np.abs(data['flattheory']), axis=0)
    assert np.max(np.abs(got['flattheory'] - data['flattheory']) / scale) < 1e-12   # the 'cubic' fixture stores the transformed vector
    for key in ['logposterior', 'loglikelihood', 'logprior']:
        ref = data[key]
        finite = np.isfinite(ref)
        assert np.array_equal(np.isneginf(ref), np.isneginf(got[key])), key
        # 'eftlike_cubic_marg': the reference's Jacobian comes from the finite-difference jacfwd shim (flatdiff is not linear in the solved
        # parameters under the cubic transform); its step / round-off error, amplified by the weakly constrained directions, is ~1e-7
        atol = 2e-6 if 'cubic_marg' in fn else 2e-9
        assert np.allclose(got[key][finite], ref[finite], rtol=1e-11, atol=atol), (key, np.max(np.abs(got[key][finite] - ref[finite])))


def test_bao_theory_fixture():
    """Oracle damped-BAO P_ell(k) and P -> xi against the reference classes' outputs (tests/golden/make_golden.py:bao_theory)."""
    from oracle import core, synthetic
    data = np.load(os.path.join(HERE, 'golden', 'bao_theory.npz'))
    X, k, s = data['X'], data['k'], data['s']
    ells = (0, 2)
    mu, wmu = core.legendre_weights(ells, 10)
    for which, kt in [('power', data['k_template_pk']), ('corr', data['k_template_xi'])]:
        pk_t, pknow_t = synthetic.pk_fid(kt), synthetic.pknow_fid(kt)
        p2c = core.PowerToCorr(s, ells)
        for x, ref in zip(X, data[which]):
            kw = dict(b1=x[2], dbeta=x[3], sigmas=x[4], sigmapar=x[5], sigmaper=x[6])
            if which == 'power':
                got = core.bao_damped_power(k, mu, wmu, kt, pk_t, pknow_t, synthetic.F_FID, x[0], x[1], **kw)
            else:
                got = p2c(core.bao_damped_power(p2c.kin, mu, wmu, kt, pk_t, pknow_t, synthetic.F_FID, x[0], x[1], **kw))
            assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref)), which


def test_kaiser_correlation_fixture():
    """Oracle Kaiser tables on kin -> get_corr (linear and cubic resampling) against the reference's KaiserTracerCorrelationFunctionMultipoles
    outputs (tests/golden/make_golden.py:kaiser_corr; reference full_shape.py:336-365, :553-574, base.py:127-136)."""
    import json
    from oracle import pipeline, synthetic
    data = np.load(os.path.join(HERE, 'golden', 'kaiser_corr.npz'))
    names, X, kt = json.loads(str(data['case']))['names'], data['X'], data['k_template']
    for order in (1, 3):
        theory = dict(kind='kaiser', s=data['s'], ells=(0, 2), nmu=8, interp_order=order, params=names[:4], defaults={},
                      template=dict(kind='shapefit', k=kt, pk_dd_fid=synthetic.pk_fid(kt), pknow_dd_fid=synthetic.pknow_fid(kt), f_fid=synthetic.F_FID,
                                    a=float(data['a']), kp=float(data['kp']), apmode='qparqper'))
        tabs = pipeline.theory_tables(theory, X, names)
        b1 = X[:, 4:5]
        got = b1**2 * tabs['pk_dd'] + 2. * b1 * tabs['pk_dt'] + tabs['pk_tt']
        ref = data['corr{:d}'.format(order)].reshape(len(X), -1)
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref)), order
    assert np.max(np.abs(data['corr3'] - data['corr1'])) > 1e-5 * np.max(np.abs(data['corr1']))   # the two resamplings are distinguishable


def test_lpt_combine_fixture():
    """Oracle velocileptors-style bias combination + cubic re-interpolation against the reference's own functions
    (tests/golden/make_golden.py:lpt_combine; reference full_shape.py:1182-1186, :1310-1312)."""
    from oracle import core
    data = np.load(os.path.join(HERE, 'golden', 'lpt_combine.npz'))
    nd = float(data['nd'])
    for pars, table, power in zip(data['pars'], data['table'], data['power']):
        assert np.allclose(core.tablevel_combine_bias_terms_poles(data['pktable'], pars, nd=nd), table, rtol=1e-13, atol=0.)
        assert np.max(np.abs(core.lpt_tracer_power(data['pktable'], data['kpt'], data['k'], pars, nd=nd) - power)) <= 1e-12 * np.max(np.abs(power))


def test_taylor_state_from_reference_fit():
    """Taylor predict layout pinned to the reference's OWN in-tree fit: tests/golden/taylor_state.npz holds center / powers / derivatives as
    produced by the unmodified ``TaylorEmulatorEngine._fit_no_operation`` (reference emulators/__init__.py:471-507; generator
    tests/golden/make_taylor_state.py) from the exact derivatives of a known polynomial; the oracle's predict on that state must return the
    polynomial, and the oracle's own term enumeration must be the same set of multi-indices in the same order."""
    from oracle import core
    data = np.load(os.path.join(HERE, 'golden', 'taylor_state.npz'))
    center, powers, derivatives = data['center'], data['powers'], data['derivatives']
    got = core.taylor_predict(data['xtest'], center, powers, derivatives)
    assert got.shape == data['ytest'].shape
    assert np.max(np.abs(got - data['ytest'])) < 1e-13 * np.max(np.abs(data['ytest']))
    assert np.array_equal(core.taylor_terms(len(center), int(powers.sum(axis=-1).max())), powers)
    # the oracle's own polynomial fit (finite differences on the stencil) lands on the same coefficients as the reference's fit of exact derivatives
    pw, dv = core.taylor_fit_polynomial(lambda X: core.taylor_predict(X, center, powers, derivatives).reshape(len(X), -1), center, len(center), 4, np.full(len(center), 0.02))
    assert np.array_equal(pw, powers)
    assert np.max(np.abs(dv.reshape(derivatives.shape) - derivatives)) < 1e-6 * np.max(np.abs(derivatives))
