"""Host logic on CPU: the C-ABI library loads and exports every symbol of include/desilike_b200.h; the plan-time folding
(desilike_b200/lowering.py) reproduces the oracle through a NumPy emulation of the kernels (tests/lowered_numpy.py)."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import configs, pipeline  # noqa: E402
import lowered_numpy  # noqa: E402


def test_library_exports_every_declared_symbol():
    import ctypes
    import desilike_b200
    from desilike_b200 import _capi, build
    build.build()
    header = open(os.path.join(ROOT, 'include', 'desilike_b200.h')).read()
    declared = set(re.findall(r'\b(dlb_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = ctypes.CDLL(build.LIBPATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib = desilike_b200.load_library()
    assert lib.dlb_abi_version() == _capi.DLB_ABI_VERSION
    assert lib.dlb_launch_count() == 0


def test_struct_layouts_match_header():
    import ctypes
    from desilike_b200 import _capi
    assert ctypes.sizeof(_capi.CoefDesc) == 8 + 4 + 4 * 4 + 4 * 4 + 4
    assert ctypes.sizeof(_capi.TermDesc) == 40
    assert _capi.PlanDesc.abi_version.offset == 0 and _capi.PlanDesc.prior_kind.offset == 8


def test_no_gpu_means_loud_failure():
    import desilike_b200
    if desilike_b200.load_library().dlb_device_count() > 0:
        pytest.skip('GPU present')
    with pytest.raises(desilike_b200.DesilikeB200Error):
        desilike_b200.Plan(configs.config_c2(nkin=20, nbins=6, order=1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'desilike_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn


def _with_interp(spec, interp):
    for obs in spec['observables']: obs['theory']['interp'] = interp
    return spec


CASES = {
    'c2': lambda: configs.config_c2(nkin=50, nbins=12, order=2),
    'c2_random': lambda: configs.config_c2(nkin=30, nbins=8, order=3, coefficients='random'),
    'c5_solved': lambda: configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True),
    'c5_dense': lambda: configs.config_c5(ntracers=2, nkin=40, nbins=10, order=2, solved=True, dense=True),
    'c5_three': lambda: configs.config_c5(ntracers=3, nkin=30, nbins=8, order=2),
}


@pytest.mark.parametrize('name', list(CASES))
def test_lowering_reproduces_oracle(name):
    from desilike_b200.lowering import Lowered
    spec = CASES[name]()
    low = Lowered(spec)
    X = configs.sample_points(spec, 24, seed=3)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    got = lowered_numpy.evaluate(low, X)
    theory = lowered_numpy.evaluate(low, X, white=False)
    assert np.max(np.abs(theory / ref['flattheory'] - 1)) < 1e-12
    for key in ['logposterior', 'loglikelihood', 'logprior']:
        assert np.allclose(got[key], ref[key], rtol=1e-10, atol=1e-8), key
    if spec['solved']:
        assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-7)


def test_marg_parameters_are_ordered_first():
    from desilike_b200.lowering import Lowered
    spec = configs.config_c5(ntracers=2, nkin=20, nbins=6, order=1, solved=True)
    for i, s in enumerate(spec['solved']): s['marg'] = bool(i % 2)
    low = Lowered(spec)
    assert low.nmarg == sum(s['marg'] for s in spec['solved'])
    flags = [spec['solved'][i]['marg'] for i in low.solved_order]
    assert flags == sorted(flags, reverse=True)
    X = configs.sample_points(spec, 6, seed=1)
    ref = pipeline.evaluate(spec, X)
    got = lowered_numpy.evaluate(low, X)
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-10, atol=1e-8)
    assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-7, atol=1e-7)


def test_unsupported_features_fail_loudly():
    from desilike_b200.lowering import Lowered, LoweringError
    spec = configs.config_kaiser_window(eftlike=False, solved=False, nkin=20, nbins=6)
    spec['observables'][0]['transform'] = 'quartic'
    with pytest.raises(LoweringError):
        Lowered(spec)


def _c3_kernel():
    spec = configs.config_c3()
    spec['observables'][0]['theory']['fftlog'] = 'kernel'
    return spec


def _with_interp(spec, interp):
    for obs in spec['observables']: obs['theory']['interp'] = interp
    return spec


THEORY_CASES = {
    'c1_kaiser': lambda: configs.config_c1(),
    'c1_kaiser_interpax': lambda: _with_interp(configs.config_c1(), 'interpax'),
    'kaiser_window_eft_marg': lambda: configs.config_kaiser_window(eftlike=True, solved=True),
    'c3_bao_fftlog': lambda: configs.config_c3(),                    # P -> xi operator folded into the term matrices (default)
    'c3_bao_fftlog_kernel': lambda: _c3_kernel(),                    # explicit FFTLog kernel payload
    'kaiser_window_cubic': lambda: configs.config_kaiser_window(eftlike=False, solved=False, transform='cubic'),   # two-stage plan
    'kaiser_window_cubic_marg': lambda: configs.config_kaiser_window(eftlike=True, solved=True, transform='cubic'),   # + Jacobian rows u'(t) dt/dx
    'c4_mlp': lambda: configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(16, 16)),
}


def _mlp_features(ig, gen, Xe):
    nh, act, d = gen.idata[:3]
    widths, dd = list(gen.idata[3:]), gen.ddata
    v = Xe * dd[:d] + dd[d:2 * d]
    off, win = 2 * d, d
    for h in widths:
        W = dd[off:off + win * h].reshape(win, h); off += win * h
        b = dd[off:off + h]; off += h
        v = v @ W + b
        v = {1: lambda t: t / (1 + np.exp(-t)), 2: lambda t: np.maximum(t, 0.), 3: np.tanh, 0: lambda t: t}[int(act)](v)
        win = h
    return np.concatenate([np.ones((len(v), 1)), np.zeros((len(v), 1)), v], axis=-1)


@pytest.mark.parametrize('name', list(THEORY_CASES))
def test_theory_payloads_reproduce_oracle(name):
    """Plan-time constants of the theory kernels (spline pivots, FFTLog tables, MLP folding) evaluated by a NumPy mirror of
    the kernels reproduce the oracle."""
    from desilike_b200.lowering import Lowered
    spec = THEORY_CASES[name]()
    low = Lowered(spec)
    X = configs.sample_points(spec, 5, seed=3)
    ref = pipeline.evaluate(spec, X, return_theory=True)
    fn = _mlp_features if name == 'c4_mlp' else lowered_numpy.theory_features
    got = lowered_numpy.evaluate(low, X, feature_fn=fn)
    theory = lowered_numpy.evaluate(low, X, feature_fn=fn, white=False)
    assert np.max(np.abs(theory - ref['flattheory']) / np.max(np.abs(ref['flattheory']), axis=0)) < 1e-11
    assert np.allclose(got['logposterior'], ref['logposterior'], rtol=1e-9, atol=1e-7)


def test_rank_reduced_whitened_problem():
    """Blocks with more data-vector columns than features are lowered onto the row space of their term matrices (QR at plan time):
    same log-likelihood, fewer columns and FLOPs; blocks with solved parameters are left alone."""
    from desilike_b200.lowering import Lowered
    spec = configs.config_c5(ntracers=2, nkin=40, nbins=40, order=2, solved=False)
    low, full = Lowered(spec), Lowered(spec, reduce_rank=False)
    assert all(b['ncols_white'] == sum(k[4] for k in b['terms']) + 1 < b['ncols'] for b in low.blocks)
    assert all(b['ncols_white'] == b['ncols'] for b in full.blocks)
    assert low.flops_per_point()['contract'] < 0.5 * full.flops_per_point()['contract']
    X = configs.sample_points(spec, 16, seed=5)
    ref = pipeline.evaluate(spec, X)
    for lw in (low, full):
        got = lowered_numpy.evaluate(lw, X)
        assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-12, atol=1e-9)
    assert np.allclose(lowered_numpy.evaluate(low, X, white=False), lowered_numpy.evaluate(full, X, white=False), rtol=0, atol=0)   # theory vectors untouched
    solved = Lowered(configs.config_c5(ntracers=2, nkin=40, nbins=40, order=2, solved=True))
    assert all(b['ncols_white'] == b['ncols'] for b in solved.blocks)


def test_spline_constants_match_scipy():
    from scipy import interpolate
    from desilike_b200 import spline
    from oracle import synthetic
    for x in [np.log10(np.geomspace(1e-3, 1., 500)), np.sort(np.random.default_rng(0).uniform(-3, 0, 40))]:
        y = synthetic.pk_fid(10**x)
        c = spline.spline_constants(x)
        M = spline.second_derivatives(c, x.size, y)
        xq = np.random.default_rng(1).uniform(x[0] - 0.05, x[-1] + 0.05, 1000)
        ref = interpolate.interp1d(x, y, kind='cubic', fill_value='extrapolate')(xq)
        assert np.max(np.abs(spline.evaluate(c, x.size, y, M, xq) / ref - 1)) < 1e-11


def test_torch_custom_op_registered_with_fake_impl():
    """torch.ops.desilike_b200.eval_batch: registered, and shape inference works without a GPU (meta tensors)."""
    import torch
    from desilike_b200 import torch_ops
    pid = torch_ops.register_plan(None, d=6, ns=3)
    outs = torch.ops.desilike_b200.eval_batch(pid, torch.empty((17, 6), dtype=torch.float64, device='meta'))
    assert [tuple(o.shape) for o in outs] == [(17,), (17,), (17,), (17, 3), (17,)]
    assert outs[4].dtype == torch.int32 and outs[0].dtype == torch.float64
    with pytest.raises(Exception):
        torch.ops.desilike_b200.eval_batch(pid, torch.empty((17, 5), dtype=torch.float64, device='meta'))
    torch_ops.unregister_plan(pid)


def test_precision_blocks_over_runs_of_observables():
    """A precision list whose blocks cover RUNS of consecutive observables (a SumLikelihood member with a joint covariance over two
    observables + a single-observable member): lowering (NumPy mirror) == oracle == the equivalent dense joint precision."""
    from desilike_b200.lowering import Lowered, LoweringError
    spec = configs.config_c5(ntracers=3, nkin=40, nbins=10, order=2, solved=True)
    P = spec['precision']
    n01 = P[0].shape[0] + P[1].shape[0]
    joint = np.zeros((n01, n01)); joint[:P[0].shape[0], :P[0].shape[0]] = P[0]; joint[P[0].shape[0]:, P[0].shape[0]:] = P[1]
    v = np.random.default_rng(6).standard_normal(n01) * np.sqrt(np.diag(joint)) * 0.1
    joint += np.outer(v, v)   # couples observables 0 and 1
    runs = dict(spec, precision=[joint, P[2]])
    n = n01 + P[2].shape[0]
    dense = np.zeros((n, n)); dense[:n01, :n01] = joint; dense[n01:, n01:] = P[2]
    X = configs.sample_points(spec, 9, seed=7)
    ref_runs, ref_dense = pipeline.evaluate(runs, X), pipeline.evaluate(dict(spec, precision=dense), X)
    assert np.allclose(ref_runs['logposterior'], ref_dense['logposterior'], rtol=1e-11, atol=1e-8)
    low = Lowered(runs)
    assert [b['ncols'] for b in low.blocks] == [n01, P[2].shape[0]]
    got = lowered_numpy.evaluate(low, X)
    assert np.allclose(got['logposterior'], ref_runs['logposterior'], rtol=1e-9, atol=1e-7)
    assert np.allclose(got['x_solved'], ref_runs['x_solved'], rtol=1e-7, atol=1e-7)
    with pytest.raises(LoweringError):
        Lowered(dict(spec, precision=[joint[:-1, :-1], P[2]]))


def test_rank_reduction_with_rank_deficient_terms():
    """The QR rank reduction stays exact when the stacked term matrices are rank deficient (a table proportional to another one, a zero table)."""
    from desilike_b200.lowering import Lowered
    spec = configs.config_c5(ntracers=1, nkin=40, nbins=40, order=2, solved=False)
    th = spec['observables'][0]['theory']
    th['tables']['pk_dt'] = 0.5 * th['tables']['pk_dd']
    th['tables']['pk_tt'] = np.zeros_like(th['tables']['pk_tt'])
    low = Lowered(spec)
    assert low.blocks[0]['ncols_white'] < low.blocks[0]['ncols']
    X = configs.sample_points(spec, 12, seed=3)
    ref, got = pipeline.evaluate(spec, X), lowered_numpy.evaluate(low, X)
    assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-12, atol=1e-9)


def test_semidefinite_precision_is_lowered():
    """A singular positive semi-definite precision (a direction projected out, as an improper-prior '.prec' marginalisation leaves it) has no
    Cholesky factor: the lowering falls back to an eigen-decomposition and reproduces the oracle's d^T P d."""
    from desilike_b200.lowering import Lowered, LoweringError
    spec = configs.config_c2(nkin=50, nbins=12, order=2)
    P = spec['precision'][0]
    v = np.random.default_rng(2).standard_normal(P.shape[0])
    Pv = P @ v
    spec['precision'] = [P - np.outer(Pv, Pv) / v.dot(Pv)]   # P v = 0
    assert abs(np.linalg.eigvalsh(spec['precision'][0])[0]) < 1e-8 * np.linalg.eigvalsh(P)[-1]
    low = Lowered(spec)
    X = configs.sample_points(spec, 10, seed=4)
    ref, got = pipeline.evaluate(spec, X), lowered_numpy.evaluate(low, X)
    assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-9, atol=1e-7)
    with pytest.raises(LoweringError):
        Lowered(dict(spec, precision=[P - 2. * np.outer(Pv, Pv) / v.dot(Pv)]))   # indefinite


def test_derived_expression_parameters_lowering():
    """Derived-expression parameters (reference parameter.py:795-808; applied per point at base.py:537) compiled to reverse-Polish programs:
    the lowered plan (NumPy mirror of the kernels) == the oracle, which evaluates the expression like the reference (Python eval); and the
    reference's own pin (tests/test_base.py:156-165): b1 = b**2 leaves the marginalised likelihood invariant."""
    from desilike_b200.lowering import Lowered, LoweringError
    from desilike_b200 import expressions
    base = configs.config_c5(ntracers=2, nkin=20, nbins=6, order=2, solved=True)
    X = configs.sample_points(base, 9, seed=4)
    ref = pipeline.evaluate(base, X)
    # T0.b1 = b**2, T1.b1 = 2 * b * exp(-c) + 0.5: the sampled columns become (qpar, qper, dm, df, b, c)
    spec = dict(base)
    names = list(base['params'])
    i0, i1 = names.index('T0.b1'), names.index('T1.b1')
    spec['params'] = [n for n in names if n not in ('T0.b1', 'T1.b1')] + ['b', 'c']
    spec['derived'] = [dict(name='T0.b1', expr='{b}**2'), dict(name='T1.b1', expr='2. * {b} * np.exp(-{c}) + 0.5')]
    spec['priors'] = dict({n: p for n, p in base['priors'].items() if n in spec['params']}, b=dict(dist='uniform', limits=(0., 3.)), c=None)
    b = np.sqrt(X[:, i0])
    c = -np.log((X[:, i1] - 0.5) / (2. * b))
    Xd = np.column_stack([X[:, [names.index(n) for n in spec['params'][:-2]]], b, c])
    orc = pipeline.evaluate(spec, Xd)
    assert np.allclose(orc['loglikelihood'], ref['loglikelihood'], rtol=1e-11, atol=1e-9)   # reparametrisation invariance of the marginalised likelihood
    low = Lowered(spec)
    assert low.d_in == len(spec['params']) and low.d == low.d_in + 2
    got = lowered_numpy.evaluate(low, Xd)
    assert np.allclose(got['loglikelihood'], orc['loglikelihood'], rtol=1e-11, atol=1e-9)
    assert np.allclose(got['logprior'], orc['logprior'], rtol=1e-12, atol=1e-12)
    assert np.allclose(got['x_solved'], orc['x_solved'], rtol=1e-9, atol=1e-9)
    with pytest.raises(LoweringError):
        Lowered(dict(spec, derived=[dict(name='T0.b1', expr='{nope}**2'), spec['derived'][1]]))
    with pytest.raises(LoweringError):
        Lowered(dict(spec, derived=[dict(name='T0.b1', expr='__import__("os").getcwd()'), spec['derived'][1]]))
    assert expressions.depends('{a} + {b} * {a}') == ['a', 'b']


@pytest.mark.parametrize('ntracers', [3, 5])
def test_split_dense_precision_lowering(ntracers):
    """Dense joint precision with more theta-dependent solved templates than one pass holds (SURVEY.md section 8d C4 'dense' stress variant:
    5 tracers x 3 counter-term rows): the split plan (un-whitened S-slots x precision-weighted y-slots, explicit entry maps, constant-row Gram as
    a plan constant) reproduces the oracle through the NumPy mirror of the kernels."""
    from desilike_b200.lowering import Lowered
    spec = configs.config_c5(ntracers=ntracers, nkin=20, nbins=6, order=2, solved=True, dense=True)
    low = Lowered(spec)
    assert low.split_dense and all(b['nslots'] <= 8 for b in low.blocks) and low.ntheory == ntracers * 18
    X = configs.sample_points(spec, 6, seed=3)
    ref = pipeline.evaluate(spec, X)
    got = lowered_numpy.evaluate(low, X)
    assert np.allclose(got['loglikelihood'], ref['loglikelihood'], rtol=1e-11, atol=1e-9)
    assert np.allclose(got['x_solved'], ref['x_solved'], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize('name', ['c4_mlp', 'c5_taylor_solved', 'eft_marg_best'])
def test_quadratic_form_tables_reproduce_the_slot_form(name):
    """Blocks with solved parameters whose terms share one feature segment can be evaluated as quadratic forms f (M_x M_y^T) f^T of the features
    (lowering.qform_tables): the NumPy mirror of that epilogue gives the same chi2 / normal equations as the slot form and the oracle; 'auto'
    only switches when it multiplies fewer columns (the full-size 5-tracer configurations do, the small ones here are forced)."""
    from desilike_b200.lowering import Lowered
    spec = {'c4_mlp': lambda: configs.config_c4(ntracers=2, nkin=40, nbins=10, hidden=(16, 16)),
            'c5_taylor_solved': lambda: configs.config_c5(ntracers=2, nkin=40, nbins=12, order=2, solved=True),
            'eft_marg_best': lambda: configs.config_kaiser_window(eftlike=True, solved=True)}[name]()
    X = configs.sample_points(spec, 7, seed=11)
    ref = pipeline.evaluate(spec, X)
    fn = _mlp_features if name == 'c4_mlp' else lowered_numpy.theory_features
    low = Lowered(spec, qform=True)
    if name == 'eft_marg_best':   # terms over different feature sub-segments (Kaiser tables): not applicable, the slot form stays
        assert all(b.get('qform') is None for b in low.blocks)
        return
    assert all(b.get('qform') is not None for b in low.blocks)
    slot = lowered_numpy.evaluate(low, X, feature_fn=fn)
    quad = lowered_numpy.evaluate(low, X, feature_fn=fn, use_qform=True)
    for key in ('logposterior', 'loglikelihood', 'x_solved'):
        assert np.allclose(quad[key], slot[key], rtol=1e-10, atol=1e-8), key
        assert np.allclose(quad[key], ref[key], rtol=1e-9, atol=1e-7), key
    assert all(b.get('qform') is None for b in Lowered(spec, qform=False).blocks)
    # full-size 5-tracer blocks: 6 terms x 66 features against 402 + 4 columns -> 'auto' picks the quadratic form (1560 instead of 2436 columns)
    if name == 'c4_mlp':
        big = Lowered(configs.config_c4(ntracers=1), with_plain=False)
        q = big.blocks[0]['qform']
        # 21 term pairs, of which (1, b1^2) and (b1, b1) of slot 0 carry the same weight b1^2 into the same Gram entry: 20 groups of 64 quadratic columns
        # (the hidden units; the constant (1, 0) head of the segment goes to direct columns: 20, the 6 r0 columns merge into them, + 6 x 4 constant rows
        # = 44 -> 6 tiles of distinct Gram entries)
        assert q is not None and q['ncolq'] == 20 * 64 + 6 * 8 and q['matrix'].shape == (64, 1328) and (q['foff'], q['K']) == (2, 64)
        for gid, (kind, _, _, _) in enumerate(q['groups']):
            if kind == 1:
                ents = [e for e, _, _ in q['dent'][gid] if e >= 0]
                assert len(ents) == len(set(ents))
