"""Plan-time constants of the not-a-knot cubic spline used by the CUDA theory kernels.

The reference interpolates the linear power spectrum with ``interp1d(..., method='cubic')`` (desilike/jax.py:211-265;
call sites full_shape.py:498, bao.py:17-19), which in the NumPy branch the reference runs here is SciPy's not-a-knot C2
cubic spline.  The tridiagonal system for the knot second derivatives has parameter-independent coefficients, so its
LU pivots are computed once here; the kernels only run the two sweeps per point (csrc/theory_kernels.cu).
"""
import numpy as np


def spline_constants(x):
    """-> flat array [x | invh | w | invb | cb | r0, r1, x0, inv_dx] (5 n + 4 doubles) for knots ``x`` (n >= 4)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 4:
        raise ValueError('not-a-knot cubic spline needs at least 4 knots')
    h = np.diff(x)
    invh = np.zeros(n); invh[:n - 1] = 1. / h
    a, b, c = np.zeros(n), np.ones(n), np.zeros(n)   # indexed by knot 1 .. n-2
    r0, r1 = h[0] / h[1], h[n - 2] / h[n - 3]
    for i in range(1, n - 1):
        a[i], b[i], c[i] = h[i - 1], 2. * (h[i - 1] + h[i]), h[i]
    b[1] += h[0] * (1. + r0); c[1] -= h[0] * r0; a[1] = 0.
    b[n - 2] += h[n - 2] * (1. + r1); a[n - 2] -= h[n - 2] * r1; c[n - 2] = 0.
    w, invb, cb = np.zeros(n), np.zeros(n), np.zeros(n)
    piv = b[1]
    invb[1], cb[1] = 1. / piv, c[1] / piv
    for i in range(2, n - 1):
        w[i] = a[i] / piv
        piv = b[i] - w[i] * c[i - 1]
        invb[i], cb[i] = 1. / piv, c[i] / piv
    tail = np.array([r0, r1, x[0], (n - 1) / (x[-1] - x[0])])
    return np.concatenate([x, invh, w, invb, cb, tail])


def is_uniform(x):
    """Knots equal x[0] + i dx to rounding (geomspace template grids in log10 k): the kernels then skip the interval search and tables."""
    x = np.asarray(x, dtype=np.float64)
    dx = (x[-1] - x[0]) / (x.size - 1)
    return bool(np.max(np.abs(x - (x[0] + dx * np.arange(x.size)))) < 1e-11 * abs(dx))


def second_derivatives(consts, n, y):
    """NumPy mirror of the device sweeps (used by the CPU tests)."""
    x, invh, w, invb, cb = (consts[i * n:(i + 1) * n] for i in range(5))
    r0, r1 = consts[5 * n], consts[5 * n + 1]
    M = np.zeros(n)
    for i in range(1, n - 1):
        M[i] = 6. * ((y[i + 1] - y[i]) * invh[i] - (y[i] - y[i - 1]) * invh[i - 1])
    carry = 0.
    for i in range(1, n - 1):
        carry = M[i] - w[i] * carry; M[i] = carry
    carry = 0.
    for i in range(n - 2, 0, -1):
        carry = M[i] * invb[i] - cb[i] * carry; M[i] = carry
    M[0] = (1. + r0) * M[1] - r0 * M[2]
    M[n - 1] = (1. + r1) * M[n - 2] - r1 * M[n - 3]
    return M


def evaluate(consts, n, y, M, xq):
    x, invh = consts[:n], consts[n:2 * n]
    xq = np.asarray(xq, dtype=np.float64)
    i = np.clip(np.searchsorted(x, xq, side='right') - 1, 0, n - 2)
    h, t = 1. / invh[i], xq - x[i]
    b = (y[i + 1] - y[i]) * invh[i] - h * (2. * M[i] + M[i + 1]) / 6.
    return y[i] + t * (b + t * (0.5 * M[i] + t * (M[i + 1] - M[i]) * invh[i] / 6.))
