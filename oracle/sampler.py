"""CPU restatement of the ensemble sampler step (TEST INFRASTRUCTURE -- checker only, never shipped).

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this.

The reference drives its vectorised log-posterior with ``emcee.EnsembleSampler(nwalkers, ndim, self.logposterior, vectorize=True)``
(reference desilike/samplers/emcee.py:68, :100-110; desilike/samplers/base.py:144-200).  emcee is a third-party package that is
absent from /root/reference and from this image, so its default move is restated from the published algorithm
(Goodman & Weare 2010 stretch move; emcee's red-blue split: the walkers are cut in two halves, each half is updated with
partners drawn from the other): ``y = c + z (x - c)``, ``z = ((a - 1) u + 1)^2 / a``, accept when
``log u' < (d - 1) log z + logp(y) - logp(x)``.

PARITY UNPINNED against emcee itself (its random stream is NumPy's Mersenne Twister; ours is counter-based Philox4x32-10 so that
results do not depend on the GPU launch geometry).  Pinned (a) bit-for-bit between this restatement and the CUDA kernels on the
same Philox stream and (b) statistically: the chain of a Gaussian target has the right mean and covariance (tests/).
"""
import numpy as np

M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011): counters / keys uint32 arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = ((p1 >> np.uint64(32)) & MASK32).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & MASK32).astype(np.uint32)
            n2 = ((p0 >> np.uint64(32)) & MASK32).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(hi, lo):
    return ((((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)).astype(np.float64)) * (1. / 9007199254740992.)


def move_randoms(seed, iteration, half, index, nc, a):
    """Stretch factor z, acceptance uniform and partner index of the walkers ``index`` (global index inside their half)."""
    index = np.asarray(index, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    code = iteration * 2 + half
    c0, c1 = (index & MASK32).astype(np.uint32), (index >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(index.shape, code & 0xFFFFFFFF, dtype=np.uint32)
    hi = (code >> 32) & 0x7FFFFFFF
    p = philox4x32_10(c0, c1, c2, np.full(index.shape, hi, dtype=np.uint32), k0, k1)
    q = philox4x32_10(c0, c1, c2, np.full(index.shape, hi | 0x80000000, dtype=np.uint32), k0, k1)
    t = (a - 1.) * u01(p[0], p[1]) + 1.
    z = t * t / a
    uacc = u01(p[2], p[3])
    partner = np.minimum((u01(q[0], q[1]) * nc).astype(np.int64), nc - 1)
    return z, uacc, partner


def stretch_step(X, logp, logposterior, seed, iteration, a=2., rank_offset=0):
    """One iteration (both halves), in place on copies; returns (X, logp, naccept).  ``logposterior(Y[n, d]) -> [n]``."""
    X, logp = np.array(X, dtype=np.float64), np.array(logp, dtype=np.float64)
    n, d = X.shape
    n0 = n // 2
    naccept = 0
    for half in range(2):
        sl, co = (slice(0, n0), slice(n0, n)) if half == 0 else (slice(n0, n), slice(0, n0))
        na, nc = sl.stop - sl.start, co.stop - co.start
        z, uacc, partner = move_randoms(seed, iteration, half, rank_offset + np.arange(na), nc, a)
        C = X[co][partner]
        Y = C + z[:, None] * (X[sl] - C)
        lpy = np.asarray(logposterior(Y), dtype=np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            lnr = (d - 1) * np.log(z) + lpy - logp[sl]
            acc = np.isfinite(lpy) & (np.log(uacc) < lnr)
        Xs, ls = X[sl], logp[sl]
        Xs[acc], ls[acc] = Y[acc], lpy[acc]
        X[sl], logp[sl] = Xs, ls
        naccept += int(acc.sum())
    return X, logp, naccept
