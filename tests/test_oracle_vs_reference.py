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
