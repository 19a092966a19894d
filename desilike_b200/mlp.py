"""Lowering of an MLP emulator engine (reference desilike/emulators/conversion.py:18-33,76-104; predict in cosmoprimo,
restated in oracle/core.py:mlp_predict) onto the CUDA path.

x-operation (min-max scale) -> affine ``v * a + b``; hidden layers run on the GPU (DMMA GEMM + activation epilogue);
the LAST dense layer and the inverse y-operations (PCA ``v @ eigvecs + mean``, min-max un-scaling) are affine in the last
hidden activations, so they are folded into each table's linear map ``table = [1, 0, h] . [c; 0; A]``.
"""
import numpy as np

from . import _capi
from ._capi import f64

ACTIVATIONS = {'identity': 0, 'silu': 1, 'relu': 2, 'tanh': 3}


def lower_mlp(gen, theory, defaults):
    mlp = theory['mlp']
    gen.kind = _capi.GEN_MLP
    d = len(gen.params)
    act = mlp.get('activation', 'silu')
    if act not in ACTIVATIONS:
        raise ValueError('activation {!r} is not on the CUDA path (supported: {})'.format(act, sorted(ACTIVATIONS)))
    if mlp.get('xlimits') is not None:
        lim = f64(mlp['xlimits'])
        a = 1. / (lim[:, 1] - lim[:, 0])
        b = -lim[:, 0] * a
    else:
        a, b = np.ones(d), np.zeros(d)
    layers = [(f64(k), f64(bb)) for k, bb in mlp['layers']]
    hidden, (W_out, b_out) = layers[:-1], layers[-1]
    widths = [k.shape[1] for k, _ in hidden]
    gen.param_default = f64([defaults.get(p, 0.) for p in gen.params])
    gen.idata = np.array([len(hidden), ACTIVATIONS[act], d] + widths, dtype=np.int32)
    gen.ddata = np.concatenate([a, b] + [np.concatenate([k.ravel(), bb]) for k, bb in hidden]) if True else None
    H = widths[-1] if widths else d
    gen.nfeat = 2 + H
    gen.const_features = {0: 1., 1: 0.}   # the (1, 0) head of the feature segment (lowering.qform_tables: no quadratic columns for constant features)
    # fold last layer + inverse y-operations:  y = h . A + c
    A, c = W_out, b_out
    if mlp.get('pca') is not None:
        eigvecs, mean = f64(mlp['pca'][0]), f64(mlp['pca'][1])
        A, c = A @ eigvecs, c @ eigvecs + mean
    if mlp.get('ylimits') is not None:
        lim = f64(mlp['ylimits'])
        scale = lim[:, 1] - lim[:, 0]
        A, c = A * scale, c * scale + lim[:, 0]
    for name, sl in theory['tables'].items():
        post = None
        if isinstance(sl, tuple): sl, post = sl   # (output indices, linear map [len(indices), nin]) applied after the selection (e.g. re-interpolation in k)
        D = np.concatenate([c[sl][None, :], np.zeros((1, c[sl].size)), A[:, sl]], axis=0)
        if post is not None: D = D @ f64(post)
        gen.tables[name] = (0, H + 2, D)


def ozaki_matmul(X, W, bias=None, slices=6, activation='identity'):
    """``act(X @ W + bias)`` through the tcgen05 int8 tensor-core path (csrc/ozaki.cu: exact integer slicing, FP64 epilogue) -- the parity entry
    of the kernels a plan uses for wide MLP hidden layers with ``Plan(..., mlp_slices=...)``.  Host arrays in, host array out."""
    import ctypes
    lib = _capi.load()
    X, W = f64(X), f64(W)
    out = np.empty((X.shape[0], W.shape[1]))
    b = f64(bias) if bias is not None else None
    _capi.check(lib.dlb_ozaki_matmul(X.ctypes.data, X.shape[0], X.shape[1], W.ctypes.data, b.ctypes.data if b is not None else None, W.shape[1], int(slices),
                                     ACTIVATIONS[activation], out.ctypes.data))
    return out


def ozaki_matmul_reference(X, W, slices=6):
    """NumPy statement of the same scheme (integer slice products are exact in float64 here): what the kernel must reproduce BIT FOR BIT."""
    def slice_rows(A):
        amax = np.max(np.abs(A), axis=1)
        E = np.where(amax > 0, np.frexp(np.where(amax > 0, amax, 1.))[1] + 1, 0)
        r = np.ldexp(A, -E[:, None])
        q = []
        for _ in range(slices):
            r = r * 128.
            qs = np.rint(r)
            r = r - qs
            q.append(qs)
        return np.array(q), E
    qa, Ea = slice_rows(np.asarray(X, dtype='f8'))
    qb, Eb = slice_rows(np.ascontiguousarray(np.asarray(W, dtype='f8').T))
    acc = []
    for t in range(slices):
        a = np.zeros((X.shape[0], W.shape[1]))
        for i in range(t + 1): a += qa[i] @ qb[t - i].T
        acc.append(a)
    # the epilogue combines the accumulators in two groups of gs = ceil(slices / 2), each EXACTLY in int64 (G < 2^46 is also exact in float64
    # here), then adds the two groups smallest first (csrc/ozaki.cu, phase A / B of the epilogue)
    gs = (slices + 1) // 2
    v = np.zeros((X.shape[0], W.shape[1]))
    for g in (1, 0):
        G = sum(acc[gs * g + u] * 2.**(7 * (gs - 1 - u)) for u in range(gs) if gs * g + u < slices)
        v = v + G * 2.**(-7 * (gs * g + gs + 1))
    return np.ldexp(np.ldexp(v, Ea[:, None]), Eb[None, :])
