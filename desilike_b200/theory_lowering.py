"""Payloads of the direct (non-emulated) theory generators: constants of the Kaiser / damped-BAO / FFTLog kernels
(csrc/theory_kernels.cu) from the theory part of a likelihood spec."""
import numpy as np
from scipy import special

from . import _capi
from ._capi import f64
from .spline import spline_constants

APMODES = {'qparqper': 0, 'qiso': 1, 'qap': 2, 'qisoqap': 3}
_AP_SLOTS = {'qparqper': ('qpar', 'qper'), 'qiso': ('qiso', None), 'qap': ('qap', None), 'qisoqap': ('qiso', 'qap')}


def weights_mu(nmu):
    """Gauss-Legendre nodes / weights on (0, 1] (reference desilike/utils.py:625-643)."""
    x, w = np.polynomial.legendre.leggauss(2 * nmu)
    return x[nmu:], (w[nmu:] + w[nmu - 1::-1]) / 2.


def legendre_weights(ells, nmu):
    """(reference desilike/theories/galaxy_clustering/base.py:201-204)."""
    mu, w = weights_mu(nmu)
    return mu, np.array([w * (2 * ell + 1) * special.legendre(ell)(mu) for ell in ells])


def _slots(gen, names_by_slot, defaults, names):
    """Fixed parameter slots of a theory kernel -> (param_idx, param_default, params list)."""
    base = {p.split('.')[-1]: p for p in gen.params}
    idx, dflt, plist = [], [], []
    for slot, default in names_by_slot:
        full = base.get(slot) if slot is not None else None
        plist.append(full if full is not None else (slot or '_'))
        idx.append(names.index(full) if (full is not None and full in names) else -1)
        dflt.append(float(defaults.get(slot, default)) if slot is not None else 0.)
    return np.array(idx, dtype=np.int32), f64(dflt), plist


def lower_theory(gen, theory, defaults, names=None):
    kind = theory['kind']
    tpl = theory['template']
    ells = tuple(theory['ells'])
    if len(ells) > 3:
        raise ValueError('at most 3 multipoles on the CUDA path')
    if int(theory.get('nmu', 8)) > 16:
        raise ValueError('at most 16 mu nodes on the CUDA path (reference defaults: 8 Kaiser, 10 BAO)')
    apmode = tpl.get('apmode', 'qparqper')
    q0, q1 = _AP_SLOTS[apmode]
    names = gen._names
    if kind == 'kaiser':
        gen.kind = _capi.GEN_KAISER
        k = f64(theory['k'])
        nmu = int(theory.get('nmu', 8))
        mu, wmu = legendre_weights(ells, nmu)
        k_t = f64(tpl['k'])
        nk, nell, nt = k.size, len(ells), k_t.size
        nin = nell * nk
        ntab = (nin + 1) & ~1
        pk = f64(tpl['pknow_dd_fid'] if tpl.get('only_now') else tpl['pk_dd_fid'])
        a, kp = float(tpl.get('a', 0.6)), float(tpl.get('kp', 0.03))
        Lt = np.log(k_t / kp)
        Tt = np.tanh(a * Lt) / a
        slots = [(q0, 1.), (q1, 1.), ('dm', 0.), ('dn', 0.), ('df', 1.), ('sigmapar', 0.), ('sigmaper', 0.), (None, 0.)]
        gen.param_idx, gen.param_default, gen.params = _slots(gen, slots, defaults, names)
        ell0 = ells.index(0) if 0 in ells else -1
        total = 2 + 4 * ntab
        base = np.concatenate([[float(tpl.get('eta', 1. / 3.)), a, float(tpl['f_fid']), 0.], k, mu, wmu.ravel(), pk, Tt, Lt,
                               spline_constants(np.log10(k_t))])
        # derived tables appended for the kernel: log10 k (so that log10(k_ap) = log10 k + log10(F(mu) / qper) costs one add per node)
        # and the spline interval widths (no division in the evaluation)
        # + the tridiagonal-sweep constants transposed to lane-major order: lane l owns the chunk of L knots starting at 1 + l L
        #   (forward) / ending at m - l L (backward); element t of every lane's chunk is contiguous -> coalesced loads
        consts = spline_constants(np.log10(k_t))
        m = nt - 2
        L = 1
        while L * 32 < m: L *= 2
        w_, invb_, cb_ = consts[2 * nt:3 * nt], consts[3 * nt:4 * nt], consts[4 * nt:5 * nt]
        wT, invbT, cbT = np.zeros((L, 32)), np.zeros((L, 32)), np.zeros((L, 32))
        for lane in range(32):
            for t in range(L):
                i = 1 + lane * L + t
                if i <= m: wT[t, lane] = w_[i]
                i = m - lane * L - t
                if i >= 1: invbT[t, lane], cbT[t, lane] = invb_[i], cb_[i]
        from .spline import is_uniform
        gen.idata = np.array([nk, nell, nmu, nt, APMODES[apmode], (ell0 << 8) | int(is_uniform(np.log10(k_t))), ntab, total, base.size, L], dtype=np.int32)
        gen.ddata = np.concatenate([base, np.log10(k), np.append(np.diff(np.log10(k_t)), 0.), wT.ravel(), invbT.ravel(), cbT.ravel()])
        gen.nfeat = total
        for t, name in enumerate(['pk_dd', 'pk_dt', 'pk_tt', 'pk11_l0']):
            gen.tables[name] = (2 + t * ntab, nin, None)
        return
    if kind in ('bao_pk', 'bao_xi'):
        from .spline import second_derivatives
        from .fftlog import fftlog_constants
        gen.kind = _capi.GEN_BAO
        xi = kind == 'bao_xi'
        nmu = int(theory.get('nmu', 10))
        mu, wmu = legendre_weights(ells, nmu)
        k_t = f64(tpl['k'])
        nell, nt = len(ells), k_t.size
        if xi:
            s_out = f64(theory['s'])
            interp_order = int(theory.get('interp_order', 1))
            if interp_order != 1:
                raise ValueError('interp_order=3 of the P -> xi resampling is not on the CUDA path (reference default is 1)')
            k = np.geomspace(1e-4, 0.6, int(300. / interp_order + 0.5)) if theory.get('kin') is None else f64(theory['kin'])  # reference base.py:67
        else:
            k = f64(theory['k'])
        nk = k.size
        consts = spline_constants(np.log10(k_t))
        pk_t, pknow_t = f64(tpl['pk_dd_fid']), f64(tpl['pknow_dd_fid'])
        M_pk, M_now = second_derivatives(consts, nt, pk_t), second_derivatives(consts, nt, pknow_t)
        from .spline import evaluate
        pknow_k = evaluate(consts, nt, pknow_t, M_now, np.log10(k))
        slots = [(q0, 1.), (q1, 1.), ('b1', 1.), ('dbeta', 1.), ('sigmas', 0.), ('sigmapar', 9.), ('sigmaper', 6.), ('df', 1.)]
        gen.param_idx, gen.param_default, gen.params = _slots(gen, slots, defaults, names)
        from .spline import is_uniform
        flags = (1 if theory.get('mode', '') == 'reciso' else 0) | (2 if xi else 0) | (4 if is_uniform(np.log10(k_t)) else 0)
        head = [float(tpl.get('eta', 1. / 3.)), float(tpl['f_fid']), float(theory.get('smoothing_radius', 15.)), 0.]
        ddata = [head, k, mu, wmu.ravel(), pknow_k, pk_t, M_pk, pknow_t, M_now, consts]
        # derived tables for the kernel (offset in idata[6]): log10 k, reconstruction smoothing S(k) = exp(-(k R)^2 / 2) ('reciso' only),
        # k^2, the WIGGLE spline (the spline is linear in its ordinates: spline(pk) - spline(pknow) = spline(pk - pknow)) and the
        # spline interval widths
        off_derived = sum(np.size(a) for a in ddata)
        rs_ = float(theory.get('smoothing_radius', 15.))
        sk = np.exp(-0.5 * (k * rs_)**2) if flags & 1 else np.zeros(nk)
        ddata += [np.log10(k), sk, k**2, pk_t - pknow_t, M_pk - M_now, np.append(np.diff(np.log10(k_t)), 0.)]
        nin = nell * nk
        fold = xi and theory.get('fftlog', 'fold') == 'fold'
        if theory.get('fftlog', 'fold') not in ('fold', 'kernel'):
            raise ValueError("theory['fftlog'] must be 'fold' (P -> xi operator folded into the term matrices) or 'kernel' (explicit FFTLog kernel)")
        if fold:
            # P_ell(kin) -> xi_ell(s) (resampling, FFTLog, interpolation onto s) is linear: fold it at plan time like the window matrix;
            # the generator is the damped-BAO kernel in P(k) mode on kin, the table 'corr' a linear map of its P_ell(kin) features
            from .fftlog import fftlog_operator
            T = fftlog_operator(k, ells, s_out)
            ns = s_out.size
            D = np.zeros((nin, nell * ns))
            for ill in range(nell): D[ill * nk:(ill + 1) * nk, ill * ns:(ill + 1) * ns] = T[ill]
            flags &= ~2
            total = 2 + nin
            header = [nk, nell, nmu, nt, APMODES[apmode], flags, off_derived, total, 0, nk, 0, 16, 2, 0, 0, 0]
            gen.idata = np.array(header, dtype=np.int32)
            gen.nfeat = total
            gen.tables['corr'] = (2, nin, D)
            gen.tables['power'] = (2, nin, None)
        elif xi:
            ns = s_out.size
            nvis = 2 + nell * ns
            poff = (nvis + 3) & ~3
            total = poff + nin
            fc = fftlog_constants(k, ells, s_out)
            off_d = sum(np.size(a) for a in ddata)
            off_d += off_d % 2    # keep the complex tables 16-byte aligned
            ddata = ddata + [np.zeros(off_d - sum(np.size(a) for a in ddata)), fc['ddata']]
            header = [nk, nell, nmu, nt, APMODES[apmode], flags, off_derived, total, ns, nk, off_d, 16, poff, fc['ng'], fc['npad_size'], fc['pad_in_left']]
            gen.idata = np.concatenate([np.array(header, dtype=np.int32), fc['idata']])
            gen.nfeat = nvis
            gen.tables['corr'] = (2, nell * ns, None)
        else:
            total = 2 + nin
            header = [nk, nell, nmu, nt, APMODES[apmode], flags, off_derived, total, 0, nk, 0, 16, 2, 0, 0, 0]
            gen.idata = np.array(header, dtype=np.int32)
            gen.nfeat = total
            gen.tables['power'] = (2, nin, None)
        gen.ddata = np.concatenate([np.ravel(a) for a in ddata])
        return
    raise ValueError('theory kind {!r} not lowered yet'.format(kind))
