"""Golden fixture pinning the Taylor emulator state layout (center / powers / derivatives) against the REFERENCE's own in-tree fit.

Runs (build container only) the unmodified ``TaylorEmulatorEngine._fit_no_operation`` (reference desilike/emulators/__init__.py:471-507) under
the stub third-party modules on a hand-built ``Samples``-like input: X = stencil points, Y = reference ``ParameterArray`` carrying ``derivs``
(reference parameter.py:204-255, :407-440) with the EXACT partial derivatives of a known vector-valued polynomial.  The resulting engine state
is stored with the keys of the reference's emulator files in tests/golden/taylor_state.npz together with test points
and the exact polynomial values; tests/test_oracle_golden.py (oracle) and tests/test_gpu_api.py (CUDA path, through EmulatedCalculator /
the 'taylor' generator) must reproduce them.
Usage: python tests/golden/make_taylor_state.py
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402


def main():
    ref_harness.setup()
    from desilike.parameter import Deriv, ParameterArray, Parameter
    from desilike.emulators import TaylorEmulatorEngine
    from desilike import mpi
    import types
    # single-process shim of the two mpi4py calls the array broadcast at the end of the fit makes (mpi4py is absent; no arithmetic involved)
    dt = types.SimpleNamespace(Commit=lambda: None, Free=lambda: None)
    mpi.MPI = types.SimpleNamespace(BYTE=types.SimpleNamespace(Create_contiguous=lambda n: dt))
    mpi.Comm.Bcast = lambda self, buf, root=0: None

    rng = np.random.default_rng(20)
    params = ['qpar', 'qper', 'dm']
    ndim, order, ny = len(params), 4, (2, 5)
    center = np.array([1., 1.02, -0.01])
    # polynomial y(x) = sum_alpha c_alpha (x - center)^alpha,  |alpha| <= order  ->  d^alpha y (center) = alpha! c_alpha
    alphas = [a for a in itertools.product(range(order + 1), repeat=ndim) if sum(a) <= order]
    coeffs = {a: rng.standard_normal(ny) * 0.5**sum(a) for a in alphas}

    def poly(X):
        d = np.asarray(X) - center
        return sum(c * np.prod(d[:, None, None, :]**np.array(a), axis=-1) for a, c in coeffs.items())

    import math
    fact = lambda a: float(np.prod([math.factorial(i) for i in a]))
    derivs = [Deriv(dict(zip(params, a))) for a in alphas]
    values = np.array([fact(a) * coeffs[a] for a in alphas])          # [nderivs, *yshape]
    # stencil X (only used by the reference to find the centre: median of the unique values per column)
    nodes = [center[j] + 0.02 * np.arange(-2, 3) for j in range(ndim)]
    X = np.array(list(itertools.product(*nodes)))
    Y = ParameterArray(np.tile(values[None], (X.shape[0],) + (1,) * values.ndim), param=Parameter('y', shape=ny), derivs=derivs)

    engine = object.__new__(TaylorEmulatorEngine)    # the cosmoprimo base class is a stub: bypass its constructor
    engine.params, engine.mpicomm = list(params), mpi.COMM_WORLD
    engine.xoperations, engine.yoperations = [], []
    engine._fit_no_operation(X, Y, None)
    assert engine.derivatives.shape[1:] == ny

    Xtest = center + rng.uniform(-0.05, 0.05, (64, ndim))
    out = dict(params=np.array(params), center=engine.center, powers=np.asarray(engine.powers), derivatives=np.asarray(engine.derivatives),
               xtest=Xtest, ytest=poly(Xtest), xshape=np.array([ndim]), yshape=np.array(ny))
    np.savez(os.path.join(ROOT, 'tests', 'golden', 'taylor_state.npz'), **out)
    print('powers', out['powers'].shape, 'derivatives', out['derivatives'].shape, 'center', out['center'])


if __name__ == '__main__':
    main()
