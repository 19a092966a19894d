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
            try: out.append(feature_fn(ig, gen, Xe, X))
            except TypeError: out.append(feature_fn(ig, gen, Xe))
    return out


def coefficients(low, X):
    C = np.ones((X.shape[0], max(len(low.coef_keys), 1)))
    for key, j in low.coef_keys.items():
        for idx, power in key:
            C[:, j] *= X[:, idx]**power
    return C


def expand(low, X):
    """Extended parameter rows: inputs + derived-expression parameters (mirror of csrc/features.cu: derived_kernel)."""
    X = np.asarray(X, dtype='f8')
    cols = [X[:, j] for j in range(X.shape[1])]
    for prog in getattr(low, 'derived_programs', []):
        st = []
        for op, arg in prog:
            if op == 0: st.append(cols[arg])
            elif op == 1: st.append(np.full(X.shape[0], low.derived_consts[arg]))
            elif op <= 6:
                b, a = st.pop(), st.pop()
                st.append({2: a + b, 3: a - b, 4: a * b, 5: a / b}.get(op) if op != 6 else np.power(a, b))
            else:
                a = st.pop()
                st.append({7: np.negative, 8: np.exp, 9: np.log, 10: np.sqrt, 11: np.log10, 12: np.abs, 13: np.sin, 14: np.cos, 15: np.tanh}[op](a))
        cols.append(st[0])
    return np.column_stack(cols)


def evaluate(low, X, feature_fn=None, white=True, use_qform=False):
    X = expand(low, X)
    N = X.shape[0]
    F, C = features(low, X, feature_fn), coefficients(low, X)
    ns = low.ns
    skind, sloc, sscale, slo, shi = low.xprior
    xprec = np.where(skind == 2, 1. / sscale**2, 0.)
    M = np.zeros((N, ns, ns)) + np.diag(xprec)
    if getattr(low, 'm0_extra', None) is not None: M = M + low.m0_extra
    gL = np.zeros((N, ns))
    chi2 = np.zeros(N)
    theory = []
    for blk in low.blocks:
        slots = np.zeros((blk['nslots'], N, blk['ncols'] if (not white or low.two_stage) else blk['ncols_white']))
        for key, mw, mp in zip(blk['terms'], blk['white'], blk['plain']):
            slot, coef, gen, foff, K = key
            c = C[:, coef] if coef >= 0 else 1.
            slots[slot] += (np.atleast_1d(c)[:, None] if coef >= 0 else 1.) * (F[gen][:, foff:foff + K] @ (mw if (white and not low.two_stage) else mp))
        if not white or low.two_stage:
            # two-stage plans (observable transform, csrc/plan.cu:run_two_stage): theory through the un-whitened matrices, transform in
            # the epilogue, then the whitening contraction
            t = slots[0] + blk['t0_plain'] if not white else sum((np.atleast_1d(C[:, key[1]])[:, None] if key[1] >= 0 else 1.) * (F[key[2]][:, key[3]:key[3] + key[4]] @ mp)
                                                                 for key, mp in zip(blk['terms'], blk['plain']) if key[0] == 0) + blk['t0_plain']
            if low.two_stage:
                d = blk['transform_data']
                with np.errstate(divide='ignore', invalid='ignore'):
                    t = np.where(d != 0., (3. * np.cbrt(t / np.where(d != 0., d, 1.)) - 2.) * d, t)
            if not white:
                theory.append(t)
                continue
            r = t @ blk['whitening_t'] - blk['r0_data_white']
            if blk['nslots'] > 1:   # solved parameters of a transformed block: Jacobian rows u'(t) dt/dx_a (csrc/contract.cu output mode, oslot > 0)
                t_plain = sum((np.atleast_1d(C[:, key[1]])[:, None] if key[1] >= 0 else 1.) * (F[key[2]][:, key[3]:key[3] + key[4]] @ mp)
                              for key, mp in zip(blk['terms'], blk['plain']) if key[0] == 0) + blk['t0_plain']
                d = blk['transform_data']
                with np.errstate(divide='ignore', invalid='ignore'):
                    up = np.where(d != 0., np.cbrt(t_plain / np.where(d != 0., d, 1.))**-2, 1.)
                chi2 += np.sum(r**2, axis=-1)
                rows = []
                for a, si in enumerate(blk['slot_solved']):
                    Sa = sum((np.atleast_1d(C[:, key[1]])[:, None] if key[1] >= 0 else 1.) * (F[key[2]][:, key[3]:key[3] + key[4]] @ mp)
                             for key, mp in zip(blk['terms'], blk['plain']) if key[0] == 1 + a)
                    rows.append((int(si), (Sa * up) @ blk['whitening_t']))
                for si, Ji in rows:
                    gL[:, si] -= np.sum(Ji * r, axis=-1)
                    for sj, Jj in rows: M[:, si, sj] += np.sum(Ji * Jj, axis=-1)
                continue
        else:
            r = slots[0] - blk['r0_white']
        if use_qform and blk.get('qform') is not None:
            # quadratic-form tables (lowering.qform_tables; csrc/contract.cu QF epilogue): the same Gram partials from f . q_matrix
            q, R, nc = blk['qform'], blk['nslots'], len(blk['const_solved'])
            f = F[q['gen']]
            out = f[:, q['foff']:q['foff'] + q['K']] @ q['matrix']
            nS = R * (R + 1) // 2
            part = np.zeros((N, nS + R * nc)) + q['const']
            for val, (ew, caw, cbw) in zip(q['wconst_val'], q['wconst_idx']):   # weighted constants (constant features)
                part[:, ew] += val * (C[:, caw] if caw >= 0 else 1.) * (C[:, cbw] if cbw >= 0 else 1.)
            for t8, (gid, k0, nv) in enumerate(q['tiles']):
                kind, e, ca, cb = q['groups'][gid]
                cols = out[:, 8 * t8:8 * t8 + 8]
                if kind == 0:
                    w = (C[:, ca] if ca >= 0 else 1.) * (C[:, cb] if cb >= 0 else 1.)
                    part[:, e] += w * np.sum(cols[:, :nv] * f[:, k0:k0 + nv], axis=-1)
                else:
                    for j in range(8):
                        ej, caj, cbj = q['dent'][gid][j]
                        if ej >= 0: part[:, ej] += (C[:, caj] if caj >= 0 else 1.) * (C[:, cbj] if cbj >= 0 else 1.) * cols[:, j]
            sidx = [None] + [int(si) for si in blk['slot_solved']]
            for a in range(R):
                for b in range(a, R):
                    v = part[:, a * R - (a * (a - 1)) // 2 + (b - a)]
                    if a == 0 and b == 0: chi2 += v
                    elif a == 0: gL[:, sidx[b]] -= v
                    else:
                        M[:, sidx[a], sidx[b]] += v
                        if a != b: M[:, sidx[b], sidx[a]] += v
            for i, ci in enumerate(blk['const_solved']):
                for a in range(R):
                    v = part[:, nS + i * R + a]
                    if a == 0: gL[:, int(ci)] -= v
                    else:
                        M[:, sidx[a], int(ci)] += v; M[:, int(ci), sidx[a]] += v
                for j, cj in enumerate(blk['const_solved']): M[:, int(ci), int(cj)] += blk['gconst_white'][i].dot(blk['gconst_white'][j])
            continue
        if blk.get('entry_kind') is not None:
            # explicit meaning of the Gram partials (split dense plans, csrc/plan.cu: build_block): pairs (a <= b) then constant rows x slots
            R, vals = blk['nslots'], []
            sl = [r] + [slots[a] for a in range(1, R)]
            for a in range(R):
                for b in range(a, R): vals.append(np.sum(sl[a] * sl[b], axis=-1))
            for i in range(len(blk['const_solved'])):
                for a in range(R): vals.append(sl[a] @ blk['gconst_white'][i])
            for e, v in enumerate(vals):
                kind, i, j = int(blk['entry_kind'][e]), int(blk['entry_i'][e]), int(blk['entry_j'][e])
                if kind == 0: chi2 += v
                elif kind == 1: gL[:, i] -= v
                elif kind == 2:
                    M[:, i, j] += v
                    if i != j: M[:, j, i] += v
            continue
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


def theory_features(ig, gen, Xe, X_full=None):
    """NumPy mirror of csrc/theory_kernels.cu for the direct theory generators, driven ONLY by the lowered payloads
    (idata / ddata) -- checks the plan-time constants (spline pivots, FFTLog tables) against the oracle on CPU."""
    from desilike_b200 import spline
    idata, dd = np.asarray(gen.idata), np.asarray(gen.ddata)
    nk, nell, nmu, nt, apmode, flags = idata[:6]
    N = Xe.shape[0]

    def ap(q0, q1, eta):
        if apmode == 0: return q0, q1
        if apmode == 1: return q0, q0
        if apmode == 2: return q0**(1 - eta), q0**(-eta)
        return q0 * q1**(1 - eta), q0 * q1**(-eta)

    if gen.kind_name == 'kaiser':
        ntab = idata[6]
        eta, f_fid = dd[0], dd[2]
        o = 4
        k = dd[o:o + nk]; o += nk
        mu = dd[o:o + nmu]; o += nmu
        wmu = dd[o:o + nell * nmu].reshape(nell, nmu); o += nell * nmu
        pk_fid, Tt, Lt = dd[o:o + nt], dd[o + nt:o + 2 * nt], dd[o + 2 * nt:o + 3 * nt]; o += 3 * nt
        consts = dd[o:o + 5 * nt + 4]
        nseg, off_seg, off_ratio, interp = idata[10], idata[11], idata[12], idata[13]
        ell0 = idata[5] >> 8
        # segment records (csrc/theory_kernels.cu: kaiser_kernel)
        segs, pos, rpos = [], off_seg, off_ratio
        for _ in range(nseg):
            flag, ncomb = idata[pos], idata[pos + 1]; pos += 2
            combs = []
            for c in range(ncomb):
                table, nf = idata[pos], idata[pos + 1]
                combs.append((table, dd[rpos], [(idata[pos + 2 + q], idata[pos + 6 + q]) for q in range(nf)])); rpos += 1; pos += 10
            segs.append((flag, combs))
        out = np.zeros((N, gen.nfeat)); out[:, 0] = 1.
        xk = consts[:nt]
        for ip in range(N):
            q0, q1, dm, dn, df, spar, sper, _ = Xe[ip]
            qpar, qper = ap(q0, q1, eta)
            y = pk_fid * np.exp(dm * Tt + dn * Lt)
            qap, jac = qpar / qper, 1. / (qpar * qper**2)
            fap = np.sqrt(1. + mu**2 * (1. / qap**2 - 1.))
            kap, muap = k[:, None] / qper * fap, np.broadcast_to(mu / qap / fap, (nk, nmu))
            if interp == 0:
                M = spline.second_derivatives(consts, nt, y)
                pint = spline.evaluate(consts, nt, y, M, np.log10(kap))
            else:
                s_ = np.diff(y) / np.diff(xk)
                fx = np.concatenate([s_[:1], 0.5 * (s_[:-1] + s_[1:]), s_[-1:]])
                xq = np.log10(kap)
                i = np.clip(np.searchsorted(xk, xq, side='right'), 1, nt - 1) - 1
                h = xk[i + 1] - xk[i]; t = (xq - xk[i]) / h
                d0, d1, dy = fx[i] * h, fx[i + 1] * h, y[i + 1] - y[i]
                pint = np.where((xq < xk[0]) | (xq > xk[-1]), np.nan, y[i] + t * (d0 + t * ((3. * dy - 2. * d0 - d1) + t * (-2. * dy + d0 + d1))))
            p = jac * pint * np.exp(-0.5 * kap**2 * (spar**2 * muap**2 + sper**2 * (1 - muap**2)))
            f = f_fid * df
            tabs = [p @ wmu.T, (f * muap**2 * p) @ wmu.T, (f**2 * muap**4 * p) @ wmu.T]   # [nk, nell]
            for iseg, (flag, combs) in enumerate(segs):
                val = np.zeros((nk, nell))
                for table, ratio, factors in combs:
                    c = ratio * np.prod([X_full[ip, idx]**power for idx, power in factors]) if factors else ratio
                    val += c * tabs[table]
                if flag & 1: val = np.tile(val[:, [ell0]], (1, nell)) if ell0 >= 0 else 0. * val
                out[ip, 2 + iseg * ntab:2 + iseg * ntab + nell * nk] = val.T.ravel()
        return out
    # BAO (+ FFTLog)
    eta, f_fid, rs = dd[0], dd[1], dd[2]
    o = 4
    k = dd[o:o + nk]; o += nk
    mu = dd[o:o + nmu]; o += nmu
    wmu = dd[o:o + nell * nmu].reshape(nell, nmu); o += nell * nmu
    pknow_k = dd[o:o + nk]; o += nk
    pk_t, M_pk, pknow_t, M_now = (dd[o + i * nt:o + (i + 1) * nt] for i in range(4)); o += 4 * nt
    consts = dd[o:o + 5 * nt + 4]
    power = np.zeros((N, nell, nk))
    for ip in range(N):
        q0, q1, b1, dbeta, sigmas, spar, sper, df = Xe[ip]
        qpar, qper = ap(q0, q1, eta)
        f = dbeta * f_fid * df
        qap = qpar / qper
        fap = np.sqrt(1. + mu**2 * (1. / qap**2 - 1.))
        kap, muap = k[:, None] / qper * fap, np.broadcast_to(mu / qap / fap, (nk, nmu))
        lk = np.log10(kap)
        pkwap = spline.evaluate(consts, nt, pk_t, M_pk, lk) - spline.evaluate(consts, nt, pknow_t, M_now, lk)
        sk = np.exp(-0.5 * (k[:, None] * rs)**2) if flags & 1 else 0.
        Cap = (b1 + f * muap**2 * (1 - sk))**2 * np.exp(-0.5 * kap**2 * (spar**2 * muap**2 + sper**2 * (1 - muap**2)))
        B = (b1 + f * mu**2 * (1 - sk))**2 / (1. + 0.5 * (sigmas * k[:, None] * mu)**2)**2
        power[ip] = ((B * pknow_k[:, None] + Cap * pkwap) @ wmu.T).T
    if not flags & 2:
        out = np.zeros((N, gen.nfeat)); out[:, 0] = 1.
        out[:, 2:] = power.reshape(N, -1)
        return out
    ns, off_d, off_i, Ng, Np, padl = idata[8], idata[10], idata[11], idata[13], idata[14], idata[15]
    fd, fi = dd[off_d:], idata[off_i:]
    c0, c1 = (fd[i * Ng:(i + 1) * Ng] for i in range(2)); o = 2 * Ng
    nu = Np // 2 + 1
    u = fd[o:o + nell * nu * 2].reshape(nell, nu, 2); o += nell * nu * 2
    u = u[..., 0] + 1j * u[..., 1]
    post = fd[o:o + nell * Ng].reshape(nell, Ng); o += nell * Ng
    o += Np   # twiddle table e^(-2 pi i j / Np), j < Np / 2
    outw = fd[o:o + nell * ns].reshape(nell, ns)
    j0, outj = fi[:Ng], fi[Ng:Ng + nell * ns].reshape(nell, ns)
    out = np.zeros((N, gen.nfeat)); out[:, 0] = 1.
    padout = Np - Ng - padl
    for ip in range(N):
        for ell in range(nell):
            pin = power[ip, ell]
            v = c0 * pin[j0] + c1 * pin[j0 + 1]   # interpolation / extrapolation as one 2-point rule, prefactor folded in
            x = np.zeros(Np); x[padl:padl + Ng] = v
            X = np.fft.fft(x)
            m = np.arange(Np)
            fac = np.where(m <= Np // 2, u[ell][np.minimum(m, Np // 2)], np.conj(u[ell][np.minimum(Np - m, Np // 2)]))
            g = np.fft.fft(X * fac).real / Np
            xi = g[padout:padout + Ng] * post[ell]
            j = outj[ell]
            out[ip, 2 + ell * ns:2 + (ell + 1) * ns] = xi[j] + outw[ell] * (xi[j + 1] - xi[j])
    return out
