"""Parity of the CUDA path (through the C-ABI) against the oracle on the same seeded inputs.  Run with ``-m gpu``.

Tolerances (BASELINE.json north_star): relative 1e-10 on theory vectors in FP64; |delta chi2| < 1e-4, i.e.
|delta logL| < 5e-5 -- the FP64 path is held to a much tighter 1e-8 * max(1, |logL|) here.

"relative 1e-10 on theory vectors" is checked twice: against the column maximum over the batch (every element, including zero crossings of
the multipoles), and ELEMENT-WISE, |got - ref| <= 1e-10 |ref|, on every element that is not within a factor 1e-3 of a zero crossing.
"""
import numpy as np
import pytest

from oracle import configs, pipeline

pytestmark = pytest.mark.gpu

THEORY_RTOL = 1e-10


def check(spec, N, seed, solved=False):
    import desilike_b200
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, N, seed=seed)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = plan.eval(X)
    theory = plan.theory(X)
    scale = np.max(np.abs(ref['flattheory']), axis=0)
    assert np.max(np.abs(theory - ref['flattheory']) / scale) < THEORY_RTOL
    away = np.abs(ref['flattheory']) > 1e-3 * scale   # element-wise, away from zero crossings
    assert away.mean() > 0.9
    assert np.max(np.abs(theory - ref['flattheory'])[away] / np.abs(ref['flattheory'])[away]) < THEORY_RTOL
    for key in ['loglikelihood', 'logprior', 'logposterior']:
        tol = 1e-8 * np.maximum(1., np.abs(ref[key]))
        assert np.all(np.abs(got[key] - ref[key]) <= tol), (key, np.max(np.abs(got[key] - ref[key])))
    assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood'])) < 5e-5  # |delta chi2| < 1e-4
    if solved:
        sc = np.maximum(1., np.abs(ref['x_solved']))
        assert np.max(np.abs(got['x_solved'] - ref['x_solved']) / sc) < 1e-7
    assert np.all(got['status'] == 0)
    plan.close()
    return got, ref


def test_c2_small(gpu_required):
    check(configs.config_c2(nkin=50, nbins=12, order=2), 300, seed=21)


def test_c2_full_shape(gpu_required):
    """BASELINE configs[1] shapes: T = 70, 3 x [3, 350] tables, W [108, 1050], n = 108."""
    check(configs.config_c2(), 1000, seed=22)


def test_c5_two_tracers_solved(gpu_required):
    check(configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True), 200, seed=23, solved=True)


def test_c5_dense_precision(gpu_required):
    check(configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True, dense=True), 130, seed=24, solved=True)


def test_c5_five_tracers(gpu_required):
    check(configs.config_c5(ntracers=5, nkin=60, nbins=40, order=3, coefficients='random'), 257, seed=25)


def test_ragged_and_edge_batches(gpu_required):
    """N = 1, N not a multiple of the point tile, out-of-prior points (-inf), device-resident torch tensors."""
    import torch
    import desilike_b200
    spec = configs.config_c2(nkin=50, nbins=12, order=2)
    plan = desilike_b200.Plan(spec)
    for N in [1, 127, 129]:
        X = configs.sample_points(spec, N, seed=N)
        ref = pipeline.evaluate(spec, X)
        got = plan.eval(X)
        assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-8)
    X = configs.sample_points(spec, 64, seed=5)
    X[3, 0] = 1.5   # qpar outside uniform[0.8, 1.2]
    X[7, 4] = -1.   # b1 outside [0, 10]
    ref = pipeline.evaluate(spec, X)
    got = plan.eval(X)
    assert np.isneginf(got['logposterior'][[3, 7]]).all() and np.isneginf(ref['logposterior'][[3, 7]]).all()
    assert (got['status'][[3, 7]] & 2).all()
    mask = np.isfinite(ref['logposterior'])
    assert np.allclose(got['logposterior'][mask], ref['logposterior'][mask], rtol=1e-9, atol=1e-8)
    Xd = torch.as_tensor(X, device='cuda')
    outd = plan.eval(Xd)
    torch.cuda.synchronize()
    assert np.array_equal(outd['logposterior'].cpu().numpy(), got['logposterior'])
    plan.close()


def test_full_size_properties(gpu_required):
    """At BASELINE's full batch size (65 536 points): chunk-independence and permutation equivariance (size-independent checks)."""
    import desilike_b200
    spec = configs.config_c2(coefficients='random')
    plan = desilike_b200.Plan(spec)
    N = 65536
    X = configs.sample_points(spec, N, seed=31)
    full = plan.eval(X)['logposterior']
    perm = np.random.default_rng(0).permutation(N)
    assert np.array_equal(plan.eval(X[perm])['logposterior'], full[perm])
    parts = np.concatenate([plan.eval(X[i:i + 10007])['logposterior'] for i in range(0, N, 10007)])
    assert np.array_equal(parts, full)
    sub = np.random.default_rng(1).choice(N, 256, replace=False)
    ref = pipeline.evaluate(spec, X[sub])['logposterior']
    assert np.allclose(full[sub], ref, rtol=1e-9, atol=1e-8)
    plan.close()


@pytest.mark.parametrize('nbins', [36, 37, 38])
def test_auxiliary_columns_across_column_block_boundary(gpu_required, nbins):
    """Marginalised tracer whose data vector (3 nbins columns) + 4 auxiliary columns (constant Jacobian rows) exactly fills (108 + 4),
    straddles (111 + 4) or starts a new 112-wide column block (114 + 4)."""
    spec = configs.config_c5(ntracers=1, nkin=60, nbins=nbins, order=2, solved=True)
    check(spec, 200, seed=51 + nbins, solved=True)


def test_fused_mlp_matches_per_layer_launches(gpu_required):
    """The fused MLP forward kernel and the per-layer launches of the contraction kernel (DLB_NO_FUSED_MLP=1) agree to rounding."""
    import os
    import desilike_b200
    spec = configs.config_c4(ntracers=2, nkin=40, nbins=10)
    X = configs.sample_points(spec, 300, seed=9)
    plan = desilike_b200.Plan(spec)
    fused = plan.eval(X)
    theory_fused = plan.theory(X)
    plan.close()
    os.environ['DLB_NO_FUSED_MLP'] = '1'
    try:
        plan = desilike_b200.Plan(spec)
        layered = plan.eval(X)
        theory_layered = plan.theory(X)
        plan.close()
    finally:
        del os.environ['DLB_NO_FUSED_MLP']
    scale = np.max(np.abs(theory_layered), axis=0)
    assert np.max(np.abs(theory_fused - theory_layered) / scale) < 1e-12
    assert np.allclose(fused['logposterior'], layered['logposterior'], rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize('name', ['c5', 'c5solved', 'c4'])
def test_five_tracer_bench_configs_full_size(gpu_required, name):
    """The 5-tracer configurations exactly as bench.py runs them (n = 2010 data points, T = 70 / MLP 64x3, 35 marginalised terms):
    every log-posterior finite, parity with the oracle on a seeded subset."""
    spec = {'c5': lambda: configs.config_c5(coefficients='random'), 'c5solved': lambda: configs.config_c5(coefficients='random', solved=True),
            'c4': lambda: configs.config_c4()}[name]()
    got, ref = check(spec, 96, seed=77, solved=bool(spec['solved']))
    assert np.all(np.isfinite(ref['logposterior'])) and np.all(np.isfinite(got['logposterior']))


def test_fp64_peak_measurement(gpu_required):
    import desilike_b200
    dmma = desilike_b200.measure_fp64_peak('dmma')
    dfma = desilike_b200.measure_fp64_peak('dfma')
    print('FP64 peak TFLOP/s: DMMA {:.2f} DFMA {:.2f}'.format(dmma, dfma))
    assert dmma > 1. and dfma > 1.


def test_posterior_hessian_and_sample_solved(gpu_required):
    """Hessian of the log-posterior in the solved parameters (dlb_eval_batch_hessian; reference likelihoods/base.py:165-187) against
    the oracle, with independent groups (block-diagonal) and one dense group; sample_solved draws with covariance (-H)^-1."""
    import desilike_b200
    for spec in [configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True), configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True, dense=True)]:
        plan = desilike_b200.Plan(spec)
        X = configs.sample_points(spec, 64, seed=3)
        got = plan.eval(X, return_hessian=True)
        ref = pipeline.evaluate(spec, X, return_hessian=True)
        scale = np.sqrt(np.abs(np.einsum('nii,njj->nij', ref['hessian'], ref['hessian'])))
        assert np.max(np.abs(got['hessian'] - ref['hessian']) / scale) < 1e-9
        assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-9)
        plan.close()
    rng = np.random.default_rng(0)
    draws = np.array([desilike_b200.sample_solved(got['x_solved'][:1], got['hessian'][:1], rng=rng)[0] for _ in range(4000)])
    cov = np.linalg.inv(-got['hessian'][0])
    assert np.allclose(draws.mean(axis=0), got['x_solved'][0], atol=5. * np.sqrt(np.diag(cov) / 4000.))
    assert np.max(np.abs(np.cov(draws.T) - cov) / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))) < 0.1


def test_pipelined_host_path_large_batch(gpu_required, monkeypatch):
    """Large host buffers go through the sub-batch pipeline (H2D | kernels | D2H on three streams; sub-batch lengths rounded to the wave quantum of
    the contraction, count from the cost model of plan.cu choose_sub_batches or forced by DLB_NSUB): same results as the one-pass device path,
    including a ragged last sub-batch."""
    import torch
    import desilike_b200
    spec = configs.config_c5(ntracers=1, nkin=40, nbins=10, order=2, solved=True)
    plan = desilike_b200.Plan(spec)
    N = 3 * 65536 + 1234
    X = configs.sample_points(spec, N, seed=77)
    dev = plan.eval(torch.as_tensor(X, device='cuda'))
    torch.cuda.synchronize()
    for forced in [None, '1', '2', '5', '8']:
        if forced is None: monkeypatch.delenv('DLB_NSUB', raising=False)
        else: monkeypatch.setenv('DLB_NSUB', forced)
        host = plan.eval(X)
        for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
            assert np.array_equal(host[key], dev[key].cpu().numpy()), (forced, key)
    plan.close()


def test_pipelined_host_path_non_resident_plan(gpu_required):
    """Plans led by a direct-theory generator (no resident A tile) keep the equal-split pipeline (from 131 072 points up, <= 4 sub-batches,
    plan.cu choose_sub_batches): host buffers == device buffers, ragged last sub-batch."""
    import torch
    import desilike_b200
    spec = configs.config_kaiser_window(nkin=200, nbins=8, eftlike=True, solved=True)   # 1202 feature columns: too wide for a resident A tile
    plan = desilike_b200.Plan(spec)
    assert plan.ns == 7
    X = configs.sample_points(spec, 2 * 65536 + 1234, seed=3)
    dev = plan.eval(torch.as_tensor(X, device='cuda'))
    torch.cuda.synchronize()
    host = plan.eval(X)
    for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
        assert np.array_equal(host[key], dev[key].cpu().numpy()), key
    plan.close()


def test_cubic_observable_transform_two_stage(gpu_required):
    """Observable transform 'cubic' (reference observables/galaxy_clustering/power_spectrum.py:400-404) on Taylor-emulated tracers: the plan runs
    theory -> transform -> whitening as two contraction launches; one tracer transformed, one not, block-diagonal and joint dense precision."""
    import desilike_b200
    for joint in [False, True]:
        spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=False)
        spec['observables'][0]['transform'] = 'cubic'
        if joint:
            P = spec['precision']
            n = sum(p.shape[0] for p in P)
            dense = np.zeros((n, n)); o = 0
            for p in P: dense[o:o + p.shape[0], o:o + p.shape[0]] = p; o += p.shape[0]
            rng = np.random.default_rng(3)
            v = rng.standard_normal(n) * np.sqrt(np.diag(dense)) * 0.1
            spec['precision'] = dense + np.outer(v, v)
        plan = desilike_b200.Plan(spec)
        X = configs.sample_points(spec, 777, seed=21)
        ref = pipeline.evaluate(spec, X, return_theory=True)
        got = plan.eval(X)
        assert np.max(np.abs(plan.theory(X) - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-10
        assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood']) / np.maximum(1., np.abs(ref['loglikelihood']))) < 1e-9
        assert np.array_equal(np.isneginf(got['logposterior']), np.isneginf(ref['logposterior']))
        plan.close()


def test_rank_reduced_plan_matches_full_plan(gpu_required):
    """Rank-reduced whitened problem (lowering.py: QR at plan time, n -> sum K_t + 1 columns) vs the full-width plan and the oracle at the
    C5 tracer size (n = 402 per tracer, block-diagonal) and for one dense joint precision."""
    import desilike_b200
    spec = configs.config_c5(ntracers=2, coefficients='random')
    X = configs.sample_points(spec, 1000, seed=31)
    ref = pipeline.evaluate(spec, X)
    red, full = desilike_b200.Plan(spec), desilike_b200.Plan(spec, reduce_rank=False)
    assert red.lowered.blocks[0]['ncols_white'] < full.lowered.blocks[0]['ncols_white'] == 402
    a, b = red.eval(X), full.eval(X)
    scale = np.maximum(1., np.abs(ref['loglikelihood']))
    assert np.max(np.abs(a['loglikelihood'] - ref['loglikelihood']) / scale) < 1e-10
    assert np.max(np.abs(a['loglikelihood'] - b['loglikelihood']) / scale) < 1e-10
    assert np.array_equal(a['status'], b['status'])
    assert np.max(np.abs(red.theory(X[:64]) - full.theory(X[:64]))) == 0.
    red.close(); full.close()
    P = spec['precision']
    n = sum(p.shape[0] for p in P)
    dense = np.zeros((n, n)); o = 0
    for p in P: dense[o:o + p.shape[0], o:o + p.shape[0]] = p; o += p.shape[0]
    v = np.random.default_rng(4).standard_normal(n) * np.sqrt(np.diag(dense)) * 0.05
    spec['precision'] = dense + np.outer(v, v)
    ref = pipeline.evaluate(spec, X[:200])
    plan = desilike_b200.Plan(spec)
    assert plan.lowered.blocks[0]['ncols_white'] < n
    got = plan.eval(X[:200])
    assert np.max(np.abs(got['loglikelihood'] - ref['loglikelihood']) / np.maximum(1., np.abs(ref['loglikelihood']))) < 1e-10
    plan.close()


def test_multi_pass_batches(gpu_required, monkeypatch):
    """Batches larger than one pass of the entry point (2^20 points, or fewer when the per-point workspace is large) are evaluated in several
    passes over the same workspace: forced here with a small pass size; host and device buffers, ragged last pass, solved parameters."""
    import torch
    import desilike_b200
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True)
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 5000, seed=41)
    one = plan.eval(X)
    monkeypatch.setenv('DLB_MAX_CHUNK', '1536')
    many = plan.eval(X)
    dev = plan.eval(torch.as_tensor(X, device='cuda'))
    torch.cuda.synchronize()
    for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
        assert np.array_equal(one[key], many[key]), key
        assert np.array_equal(one[key], dev[key].cpu().numpy()), key
    plan.close()


def test_per_point_failures_status_bits(gpu_required):
    """Per-point failure paths of the C-ABI (include/desilike_b200.h status bits): a non-finite theory (NONFINITE -> NaN), a singular
    normal matrix of the solved parameters (NOT_SPD -> NaN, x_solved NaN), both isolated to their points."""
    import desilike_b200
    from desilike_b200 import _capi
    spec = configs.config_c2(nkin=50, nbins=12, order=2)
    spec['priors']['b1'] = None   # no limits: a huge bias is inside the prior and overflows the theory
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 70, seed=5)
    ib1 = spec['params'].index('b1')
    X[11, ib1] = 1e200            # b1^2 = inf
    X[12, ib1] = np.nan
    got = plan.eval(X)
    ref = pipeline.evaluate(spec, X)
    assert got['status'][11] & _capi.PT_NONFINITE and np.isnan(got['loglikelihood'][11]) and np.isnan(got['logposterior'][11])
    assert got['status'][12] & _capi.PT_NONFINITE and not np.isfinite(got['logposterior'][12])
    ok = np.ones(70, dtype=bool); ok[[11, 12]] = False
    assert np.all(got['status'][ok] == 0)
    assert np.allclose(got['logposterior'][ok], ref['logposterior'][ok], rtol=1e-9, atol=1e-8)
    with np.errstate(all='ignore'):
        assert not np.isfinite(ref['logposterior'][11])
    plan.close()
    # singular normal equations: a solved parameter without prior whose template vanishes identically
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True)
    name = spec['solved'][3]['name']
    spec['solved'][3]['prior'] = None
    for obs in spec['observables']:
        for term in obs['terms']:
            if term.get('solved') == name: term['colvec'] = np.zeros_like(term['colvec'])
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 50, seed=6)
    got = plan.eval(X)
    assert np.all(got['status'] & _capi.PT_NOT_SPD) and np.all(np.isnan(got['logposterior'])) and np.all(np.isnan(got['x_solved']))
    plan.close()


def test_dense_precision_many_solved_templates_split(gpu_required):
    """SURVEY.md section 8d C4 / C5 'dense' stress variants: 5 tracers, joint dense precision, 15 theta-dependent + 20 constant marginalised
    templates in ONE likelihood block (16 accumulator slots > the 8 of one pass): the split plan == the oracle, Taylor and MLP theories."""
    check(configs.config_c5(ntracers=5, nkin=40, nbins=10, order=2, solved=True, dense=True), 150, seed=31, solved=True)
    check(configs.config_c4(ntracers=5, nkin=40, nbins=10, hidden=(16, 16), dense=True), 130, seed=32, solved=True)
    import desilike_b200
    spec = configs.config_c5(ntracers=4, nkin=40, nbins=10, order=2, solved=True, dense=True)
    for s in spec['solved'][::2]: s['marg'] = False     # mixed '.marg' / '.best'
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, 64, seed=33)
    got, ref = plan.eval(X, return_hessian=True), pipeline.evaluate(spec, X, return_hessian=True)
    assert np.all(np.abs(got['logposterior'] - ref['logposterior']) <= 1e-8 * np.maximum(1., np.abs(ref['logposterior'])))
    assert np.allclose(got['hessian'], ref['hessian'], rtol=1e-8, atol=1e-10 * np.max(np.abs(ref['hessian'])))
    plan.close()


def test_cubic_transform_with_solved_parameters(gpu_required):
    """transform='cubic' together with analytically marginalised terms: flatdiff = u(t(x)) - data is not linear in x, the reference linearises
    at the defaults (jacfwd through the transform, likelihoods/base.py:152-158).  Stage 1 of the plan writes [u | u'(t) dt/dx_i] (constant templates
    become slots too), stage 2 whitens all of them; direct Kaiser theory (7 templates) and Taylor-emulated tracers, '.marg' and '.best'."""
    check(configs.config_kaiser_window(eftlike=True, solved=True, transform='cubic'), 300, seed=41, solved=True)
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True)
    spec['observables'][1]['transform'] = 'cubic'
    spec['observables'][0]['transform'] = 'cubic'
    for s in spec['solved'][1::2]: s['marg'] = False
    check(spec, 257, seed=42, solved=True)
