"""FP64-grade GEMM on the tcgen05 int8 tensor cores (csrc/ozaki.cu): the kernels reproduce the NumPy statement of the exact-slicing scheme bit for
bit, and FP64 matmul to the accuracy the slice count promises; wide-MLP plans using them stay inside |delta chi2| < 1e-4 (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import configs, pipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(300, 512, 512), (129, 200, 70), (1, 64, 16), (1000, 1024, 192)])
def test_ozaki_matmul_bit_exact_and_accurate(gpu_required, shape):
    from desilike_b200 import mlp
    rows, K, N = shape
    rng = np.random.default_rng(rows)
    X = rng.standard_normal((rows, K)) * np.exp(rng.uniform(-3., 3., (rows, 1)))   # rows of very different scales: per-row exponents
    W = rng.standard_normal((K, N)) * np.sqrt(2. / K)
    X[0] = 0.                                                                       # an all-zero row
    ref = X @ W
    for slices in (4, 6):
        got = mlp.ozaki_matmul(X, W, slices=slices)
        emu = mlp.ozaki_matmul_reference(X, W, slices=slices)
        assert np.array_equal(got, emu), np.max(np.abs(got - emu))
        scale = np.max(np.abs(X), axis=1)[:, None] * np.max(np.abs(W), axis=0)[None, :] * K
        # dropped slice products i + j >= slices and the operands' tails: <= 16 (slices + 2) 2^(-7 (slices + 1)) relative to max|row| max|col| K
        assert np.max(np.abs(got - ref) / np.maximum(scale, 1e-300)) < 16. * (slices + 2) * 2.**(-7 * (slices + 1))
    bias = rng.standard_normal(N)
    got = mlp.ozaki_matmul(X, W, bias=bias, slices=6, activation='silu')
    v = ref + bias
    assert np.allclose(got, v / (1. + np.exp(-v)), rtol=1e-9, atol=1e-10 * np.max(np.abs(v)))
