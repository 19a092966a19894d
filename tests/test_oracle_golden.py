"""The oracle against the committed golden fixtures (outputs of the reference's own NumPy path, tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import configs, pipeline

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, 'golden', '*.npz')))


def load_case(fn):
    data = np.load(fn)
    case = json.loads(str(data['case']))
    spec = getattr(configs, case['config'])(**case['kwargs'])
    if case['mode'] == '.best':
        for s in spec['solved']: s['marg'] = False
    return case, spec, data


def test_fixtures_present():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize('fn', FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_oracle_matches_reference_fixture(fn):
    case, spec, data = load_case(fn)
    X = data['X']
    got = pipeline.evaluate(spec, X, return_theory=True)
    scale = np.max(np.abs(data['flattheory']), axis=0)
    if spec['observables'][0].get('transform') is None:
        assert np.max(np.abs(got['flattheory'] - data['flattheory']) / scale) < 1e-12
    for key in ['logposterior', 'loglikelihood', 'logprior']:
        ref = data[key]
        finite = np.isfinite(ref)
        assert np.array_equal(np.isneginf(ref), np.isneginf(got[key])), key
        assert np.allclose(got[key][finite], ref[finite], rtol=1e-11, atol=2e-9), (key, np.max(np.abs(got[key][finite] - ref[finite])))
