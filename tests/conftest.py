import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def has_gpu():
    try:
        import desilike_b200
        return desilike_b200.load_library().dlb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope='session')
def gpu_required():
    """GPU tests must FAIL (not skip) when the CUDA path is unusable on a GPU box: no silent fallback."""
    import desilike_b200
    lib = desilike_b200.load_library()
    assert lib.dlb_device_count() > 0, 'no CUDA device visible'
    return lib
