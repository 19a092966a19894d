"""Randomised host-logic check: the lowered plan (NumPy mirror of the kernels, tests/lowered_numpy.py) against the oracle over seeded random small
configurations and lowering options (rank reduction, term merging, quadratic-form tables on / off).  CPU only."""
import numpy as np
import pytest

import lowered_numpy
from oracle import configs, pipeline
from desilike_b200.lowering import Lowered, LoweringError


def _random_case(rng):
    kind = rng.choice(['c5', 'c4', 'c2', 'kaiser'])
    seed = int(rng.integers(1, 1000))
    if kind == 'c5':
        kw = dict(seed=seed, ntracers=int(rng.integers(1, 4)), order=int(rng.integers(1, 4)), nkin=int(rng.integers(12, 50)), nbins=int(rng.integers(3, 14)),
                  solved=bool(rng.integers(0, 2)), dense=bool(rng.integers(0, 2)))
        return kind, kw, configs.config_c5(**kw), None
    if kind == 'c4':
        from test_capi_and_lowering import _mlp_features
        h = int(rng.choice([4, 8, 12, 16, 20]))
        kw = dict(seed=seed, ntracers=int(rng.integers(1, 3)), hidden=(h,) * int(rng.integers(1, 3)), nkin=int(rng.integers(12, 40)), nbins=int(rng.integers(3, 12)),
                  solved=bool(rng.integers(0, 2)), dense=bool(rng.integers(0, 2)))
        return kind, kw, configs.config_c4(**kw), _mlp_features
    if kind == 'c2':
        kw = dict(seed=seed, order=int(rng.integers(1, 4)), nkin=int(rng.integers(12, 50)), nbins=int(rng.integers(3, 14)))
        return kind, kw, configs.config_c2(**kw), None
    kw = dict(seed=seed, nkin=int(rng.integers(20, 50)), nbins=int(rng.integers(4, 12)), eftlike=bool(rng.integers(0, 2)), solved=bool(rng.integers(0, 2)),
              nmu=int(rng.choice([6, 8, 10])), systematic=bool(rng.integers(0, 2)))
    if not kw['eftlike'] and not kw['systematic']: kw['solved'] = False
    return kind, kw, configs.config_kaiser_window(**kw), lowered_numpy.theory_features


@pytest.mark.parametrize('seed', [11, 12, 13, 14, 15, 16])
def test_random_configurations_lowered_plan_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    for _ in range(3):
        kind, kw, spec, feature_fn = _random_case(rng)
        X = configs.sample_points(spec, 7, seed=int(rng.integers(1, 1000)))
        ref = pipeline.evaluate(spec, X)
        fin = np.isfinite(ref['logposterior'])
        for opts in [dict(), dict(reduce_rank=False), dict(merge_terms=False), dict(qform=True), dict(qform=False)]:
            try:
                low = Lowered(spec, **opts)
            except LoweringError:
                if opts.get('qform') is True: continue   # (a forced quadratic form may be inapplicable to the configuration)
                raise
            got = lowered_numpy.evaluate(low, X, feature_fn=feature_fn, use_qform=any(b.get('qform') for b in low.blocks))
            assert np.allclose(got['logposterior'][fin], ref['logposterior'][fin], rtol=1e-8, atol=1e-6), (kind, kw, opts)
            if np.size(ref.get('x_solved', ())):
                scale = max(1., np.max(np.abs(ref['x_solved'][fin])))
                assert np.allclose(got['x_solved'][fin], ref['x_solved'][fin], rtol=1e-6, atol=1e-7 * scale), (kind, kw, opts)
