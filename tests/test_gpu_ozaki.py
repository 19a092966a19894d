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
    for slices in (4, 5, 6, 8):
        got = mlp.ozaki_matmul(X, W, slices=slices)
        emu = mlp.ozaki_matmul_reference(X, W, slices=slices)
        assert np.array_equal(got, emu), np.max(np.abs(got - emu))
        scale = np.max(np.abs(X), axis=1)[:, None] * np.max(np.abs(W), axis=0)[None, :] * K
        # dropped slice products i + j >= slices and the operands' tails: <= 16 (slices + 2) 2^(-7 (slices + 1)) relative to max|row| max|col| K
        # (+ the float64 rounding of the result itself and of the reference product: 8 slices reach that floor)
        assert np.max(np.abs(got - ref) / np.maximum(scale, 1e-300)) < 16. * (slices + 2) * 2.**(-7 * (slices + 1)) + 2.**-50
    bias = rng.standard_normal(N)
    got = mlp.ozaki_matmul(X, W, bias=bias, slices=6, activation='silu')
    v = ref + bias
    assert np.allclose(got, v / (1. + np.exp(-v)), rtol=1e-9, atol=1e-10 * np.max(np.abs(v)))


def test_wide_mlp_plan_on_int8_tensor_cores(gpu_required):
    """Wide MLP emulators with ``Plan(..., mlp_slices=6)``: hidden layers with fan-in >= 64 run on the tcgen05 int8 path; the likelihood stays within
    |delta chi2| < 1e-4 of the oracle (BASELINE.json north_star tolerance for tensor-core arithmetic) -- in fact ~1e-7 -- and the marginalised
    parameters agree; 4 slices (28 bits) visibly do not reach the FP64 path's 1e-8 but the plan still evaluates."""
    import desilike_b200
    spec = configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(128, 96, 128))
    X = configs.sample_points(spec, 333, seed=51)
    ref = pipeline.evaluate(spec, X)
    plan = desilike_b200.Plan(spec, mlp_slices=6)
    # fan-in 128 / 96 hidden layers AND the likelihood blocks (their terms all contract the last 128-wide hidden layer) run on the int8 path
    assert plan.kernels['int8_layers'] and plan.kernels['int8_contraction']
    got = plan.eval(X)
    assert 2. * np.max(np.abs(got['loglikelihood'] - ref['loglikelihood'])) < 1e-4
    assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood'])) < 1e-6 * np.max(np.abs(ref['loglikelihood']))
    assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-5, atol=1e-5)
    fp64_plan = desilike_b200.Plan(spec)
    assert not fp64_plan.kernels['int8_layers'] and not fp64_plan.kernels['int8_contraction']
    fp64 = fp64_plan.eval(X)
    assert np.max(np.abs(got['loglikelihood'] - fp64['loglikelihood'])) < 1e-6 * np.max(np.abs(ref['loglikelihood']))
    plan.close()
    coarse = desilike_b200.Plan(spec, mlp_slices=3).eval(X)
    assert np.all(np.isfinite(coarse['logposterior'])) and np.max(np.abs(coarse['loglikelihood'] - ref['loglikelihood'])) > 1e-7
    with pytest.raises(desilike_b200.DesilikeB200Error):
        desilike_b200.Plan(spec, mlp_slices=9)


def test_wide_mlp_contraction_at_the_stress_variant_shapes(gpu_required):
    """The likelihood blocks of the `c4wide` bench workload (K = 514 features, 6 terms x (402 + 4) columns per tracer, 4 slots) as ONE int8 GEMM per
    tracer + ozaki_gram_kernel: |delta chi2| < 1e-4 against the oracle, solved parameters equal, and the same numbers as the FP64 DMMA plan."""
    import desilike_b200
    spec = configs.config_c4(ntracers=2, hidden=(512, 512))
    X = configs.sample_points(spec, 257, seed=7)
    ref = pipeline.evaluate(spec, X)
    plan = desilike_b200.Plan(spec, mlp_slices=6)
    assert plan.kernels['int8_layers'] and plan.kernels['int8_contraction']
    got = plan.eval(X)
    fp64 = desilike_b200.Plan(spec).eval(X)
    for other in (ref, fp64):
        assert 2. * np.max(np.abs(got['loglikelihood'] - other['loglikelihood'])) < 1e-4
        assert np.allclose(got['x_solved'], other['x_solved'], rtol=1e-5, atol=1e-5)
    assert np.array_equal(got['status'], fp64['status'])


@pytest.mark.parametrize('cluster', [2, 4])
def test_ozaki_matmul_cluster_multicast_path(gpu_required, monkeypatch, cluster):
    """Opt-in thread-block clusters with TMA multicast of the A operand (DLB_OZAKI_CLUSTER): same bits as the default launch and as the NumPy statement,
    including column-tile counts that are not a multiple of the cluster size (dummy tiles)."""
    from desilike_b200 import mlp
    rng = np.random.default_rng(cluster)
    for rows, K, N in [(500, 512, 512), (333, 200, 6 * 64 + 10), (130, 96, 5 * 64)]:
        X, W = rng.standard_normal((rows, K)), rng.standard_normal((K, N))
        base = mlp.ozaki_matmul(X, W, slices=6)
        monkeypatch.setenv('DLB_OZAKI_CLUSTER', str(cluster))
        got = mlp.ozaki_matmul(X, W, slices=6)
        monkeypatch.delenv('DLB_OZAKI_CLUSTER')
        assert np.array_equal(got, base) and np.array_equal(got, mlp.ozaki_matmul_reference(X, W, slices=6))
