"""Quadratic-form evaluation of blocks with solved parameters (lowering.qform_tables + the QF epilogue of csrc/contract.cu): same chi2, marginalised /
profiled parameters and status as the slot form and as the oracle."""
import numpy as np
import pytest

from oracle import configs, pipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['c4_small_forced', 'c5solved_small_forced', 'c4_full_auto', 'c5solved_full_auto'])
def test_quadratic_form_plan_matches_slot_form_and_oracle(gpu_required, name):
    import desilike_b200
    spec, force = {'c4_small_forced': (lambda: configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(16, 16)), True),
                   'c5solved_small_forced': (lambda: configs.config_c5(ntracers=2, nkin=40, nbins=12, order=2, solved=True), True),
                   'c4_full_auto': (lambda: configs.config_c4(ntracers=2), 'auto'),
                   'c5solved_full_auto': (lambda: configs.config_c5(ntracers=2, solved=True), 'auto')}[name]
    spec = spec()
    X = configs.sample_points(spec, 1000 if 'small' in name else 300, seed=13)
    ref = pipeline.evaluate(spec, X)
    quad = desilike_b200.Plan(spec, qform=force)
    slot = desilike_b200.Plan(spec, qform=False)
    assert quad.kernels['quadratic_form'] and not slot.kernels['quadratic_form']
    got, base = quad.eval(X), slot.eval(X)
    scale = np.max(np.abs(ref['loglikelihood']))
    for other in (ref, base):
        assert np.max(np.abs(got['loglikelihood'] - other['loglikelihood'])) < 1e-10 * scale
        assert np.allclose(got['x_solved'], other['x_solved'], rtol=1e-7, atol=1e-8 * np.max(np.abs(ref['x_solved'])))
        assert np.allclose(got['logposterior'], other['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.array_equal(got['status'], base['status'])
    # theory vectors still come from the term matrices
    assert np.array_equal(quad.theory(X[:50]), slot.theory(X[:50]))
    # several passes / ragged tails
    sub = quad.eval(X[:77])
    assert np.array_equal(sub['logposterior'], got['logposterior'][:77])


def test_quadratic_form_plan_hessian_and_multipass(gpu_required, monkeypatch):
    """The Hessian entry and the several-passes path (DLB_MAX_CHUNK) run the same quadratic-form launch."""
    import desilike_b200
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=12, order=2, solved=True)
    X = configs.sample_points(spec, 2000, seed=17)
    quad, slot = desilike_b200.Plan(spec, qform=True), desilike_b200.Plan(spec, qform=False)
    a, b = quad.eval(X, return_hessian=True), slot.eval(X, return_hessian=True)
    assert np.allclose(a['hessian'], b['hessian'], rtol=1e-9, atol=1e-9 * np.max(np.abs(b['hessian'])))
    assert np.allclose(a['logposterior'], b['logposterior'], rtol=1e-10, atol=1e-8)
    one = quad.eval(X)   # (same finish kernel as the multi-pass run below; the Hessian entry uses the warp-per-point one)
    monkeypatch.setenv('DLB_MAX_CHUNK', '768')
    c = quad.eval(X)
    assert np.array_equal(c['logposterior'], one['logposterior']) and np.array_equal(c['x_solved'], one['x_solved'])


@pytest.mark.parametrize('name', ['mlp', 'taylor'])
def test_quadratic_form_non_finite_point_stays_isolated(gpu_required, name):
    """A NaN / overflowing parameter point poisons its own result only: the feature tiles the epilogue reads never cross into a neighbouring
    point's row of the resident A tile (points at the edges of the 64 / 128-point tiles included)."""
    import desilike_b200
    from desilike_b200 import _capi
    spec = (configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(20, 20)) if name == 'mlp'
            else configs.config_c5(ntracers=2, nkin=40, nbins=12, order=2, solved=True))
    for key in spec['priors']: spec['priors'][key] = None
    plan = desilike_b200.Plan(spec, qform=True)
    assert plan.kernels['quadratic_form']
    X = configs.sample_points(spec, 700, seed=23)
    clean = plan.eval(X)
    bad = [0, 63, 64, 127, 128, 255, 699]
    Y = X.copy()
    Y[bad[::2], 0] = np.nan
    Y[bad[1::2], :] = 1e200
    got = plan.eval(Y)
    ok = np.ones(len(X), dtype=bool); ok[bad] = False
    assert not np.any(np.isfinite(got['logposterior'][bad[::2]])) and np.all(got['status'][bad[::2]] & (_capi.PT_NONFINITE | _capi.PT_NOT_SPD))
    for key in ['logposterior', 'loglikelihood', 'x_solved', 'status']:
        assert np.array_equal(got[key][ok], clean[key][ok]), key
    plan.close()
