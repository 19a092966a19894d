"""NumPy emulation of what the CUDA kernels do with a *lowered* plan (TEST HELPER, CPU-only host-logic checks).

It lets the ``-m "not gpu"`` suite verify the plan-time folding of desilike_b200/lowering.py against the oracle without
a GPU.  It mirrors csrc/features.cu (Taylor features, coefficients), csrc/contract.cu (term contraction + Gram
partials) and csrc/solve.cu (normal equations) at the level of the arrays handed to dlb_plan_create.
"""
import numpy as np


def features(low, X, feature_fn=None):
    """-> list of per-generator feature arrays [N, nfeat]."""
    out = []
    for ig, gen in enumerate(low.gens):
        cols = []
        for j, p in enumerate(gen.params):
            idx = int(gen.param_idx[j])
            cols.append(X[:, idx] if idx >= 0 else np.full(X.shape[0], gen.param_default[j]))
        Xe = np.stack(cols, axis=-1) if cols else np.zeros((X.shape[0], 0))
        if gen.kind_name == 'taylor':
            T = gen.nfeat - 2
            powers = np.asarray(gen.idata).reshape(T, len(gen.params))
            mono = np.prod((Xe - gen.ddata)[:, None, :]**powers[None], axis=-1)
            out.append(np.concatenate([np.ones((X.shape[0], 1)), np.zeros((X.shape[0], 1)), mono], axis=-1))
        else:
            out.append(feature_fn(ig, gen, Xe))
    return out


def coefficients(low, X):
    C = np.ones((X.shape[0], max(len(low.coef_keys), 1)))
    for key, j in low.coef_keys.items():
        for idx, power in key:
            C[:, j] *= X[:, idx]**power
    return C


def evaluate(low, X, feature_fn=None, white=True):
    X = np.asarray(X, dtype='f8')
    N = X.shape[0]
    F, C = features(low, X, feature_fn), coefficients(low, X)
    ns = low.ns
    skind, sloc, sscale, slo, shi = low.xprior
    xprec = np.where(skind == 2, 1. / sscale**2, 0.)
    M = np.zeros((N, ns, ns)) + np.diag(xprec)
    gL = np.zeros((N, ns))
    chi2 = np.zeros(N)
    theory = []
    for blk in low.blocks:
        slots = np.zeros((blk['nslots'], N, blk['ncols']))
        for key, mw, mp in zip(blk['terms'], blk['white'], blk['plain']):
            slot, coef, gen, foff, K = key
            c = C[:, coef] if coef >= 0 else 1.
            slots[slot] += (np.atleast_1d(c)[:, None] if coef >= 0 else 1.) * (F[gen][:, foff:foff + K] @ (mw if white else mp))
        if not white:
            theory.append(slots[0] + blk['t0_plain'])
            continue
        r = slots[0] - blk['r0_white']
        chi2 += np.sum(r**2, axis=-1)
        rows = [(int(si), slots[1 + a]) for a, si in enumerate(blk['slot_solved'])]
        rows += [(int(si), np.broadcast_to(blk['gconst_white'][i], (N, blk['ncols']))) for i, si in enumerate(blk['const_solved'])]
        for si, Ji in rows:
            gL[:, si] -= np.sum(Ji * r, axis=-1)
            for sj, Jj in rows:
                M[:, si, sj] += np.sum(Ji * Jj, axis=-1)
    if not white:
        return np.concatenate(theory, axis=-1)
    kind, lo, hi, loc, scale = low.prior
    lp = np.zeros(N)
    for j in range(low.d):
        x = X[:, j]
        isin = (lo[j] <= x) & (x <= hi[j])
        val = -0.5 * ((x - loc[j]) / scale[j])**2 if kind[j] == 2 else 0.
        lp += np.where(isin, val, -np.inf)
    ll = -0.5 * chi2
    xs = np.zeros((N, ns))
    if ns:
        g0 = -(low.x0 - sloc) * xprec
        for ip in range(N):
            g = g0 + gL[ip]
            Lc = np.linalg.cholesky(M[ip])
            dx = np.linalg.solve(M[ip], g)
            x = low.x0 + dx
            ll[ip] += 0.5 * (-g.dot(dx) + np.sum(xprec * dx**2)) + gL[ip].dot(dx)
            ll[ip] -= 0.5 * 2. * np.sum(np.log(np.diag(Lc)[:low.nmarg]))
            for i in range(ns):
                isin = slo[i] <= x[i] <= shi[i]
                lp[ip] += (-0.5 * (x[i] - sloc[i])**2 * xprec[i] if skind[i] == 2 else 0.) if isin else -np.inf
            xs[ip] = x
    return dict(loglikelihood=ll, logprior=lp, logposterior=ll + lp, x_solved=xs[:, np.argsort(low.solved_order)] if ns else xs)
