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


KAISER_TABLE_ID = {'pk_dd': 0, 'pk_dt': 1, 'pk_tt': 2, 'pk11_l0': 0}


def kaiser_set_segments(gen, segments):
    """(Re)build the payload of a Kaiser generator for the output segments ``segments`` = [dict(name, flag, combs=[(table id, ratio, ((param column,
    power), ...)), ...])]: segment = sum_combs ratio * prod theta^power * table (csrc/theory_kernels.cu: kaiser_kernel).  Sets gen.tables /
    gen.nfeat / gen.idata / gen.ddata."""
    kz = gen._kaiser
    if not 1 <= len(segments) <= 8:
        raise ValueError('a Kaiser generator holds 1 to 8 output segments')
    nseg, ntab, nin = len(segments), kz['ntab'], kz['nin']
    total = 2 + nseg * ntab
    recs, ratios = [], []
    for seg in segments:
        recs += [int(seg.get('flag', 0)), len(seg['combs'])]
        for table, ratio, factors in seg['combs']:
            factors = list(factors)
            if len(factors) > 4: raise ValueError('too many factors in a merged bias monomial')
            recs += [int(table), len(factors)] + [int(i) for i, _ in factors] + [0] * (4 - len(factors)) + [int(p) for _, p in factors] + [0] * (4 - len(factors))
            ratios.append(float(ratio))
    base, derived = kz['base'], kz['derived']   # derived ends with the 6 converged-pivot constants the kernel reads just before the ratios
    header = kz['head'] + [total, base.size, 0, nseg, 16, base.size + derived.size, kz['interp'], kz['pad'], kz['nw_max']]
    gen.idata = np.array(header + recs, dtype=np.int32)
    gen.ddata = np.concatenate([base, derived, ratios])
    gen.nfeat = total
    gen.tables = {seg['name']: (2 + i * ntab, nin, kz.get('fold')) for i, seg in enumerate(segments)}   # fold: P_ell(kin) -> xi_ell(s) operator or None
    gen.segments = list(segments)


def _converged_pivots(consts, nt, uniform):
    """Uniform knots: away from the two ends the pivots (w, 1 / b, c / b) of the not-a-knot tridiagonal solve converge geometrically to
    constants.  -> [w_inf, invb_inf, cb_inf, first knot, last knot, 0] such that on [first, last] every tabulated pivot equals the constant to
    1e-15 relative (the kernel then runs its sweeps without table loads); first > last when there is no such range."""
    w, invb, cb = consts[2 * nt:3 * nt], consts[3 * nt:4 * nt], consts[4 * nt:5 * nt]
    if not uniform or nt < 64: return np.array([0., 0., 0., 1., 0., 0.])
    mid = nt // 2
    ref = np.array([w[mid], invb[mid], cb[mid]])
    ok = np.all(np.abs(np.stack([w, invb, cb], axis=-1) - ref) <= 1e-15 * np.abs(ref), axis=-1)
    ok[0] = ok[-1] = False
    lo = mid
    while lo - 1 >= 1 and ok[lo - 1]: lo -= 1
    hi = mid
    while hi + 1 <= nt - 2 and ok[hi + 1]: hi += 1
    return np.array([ref[0], ref[1], ref[2], float(lo), float(hi), 0.])


def _ap_log_span(apmode, q0, q1, gen, defaults, priors, eta, mu):
    """[min, max] of log10(F(mu) / qper) over the prior box of the AP parameters, or None when it is unbounded."""
    base = {p.split('.')[-1]: p for p in gen.params}

    def limits(slot):
        if slot is None: return (1., 1.)
        full = base.get(slot)
        if full is None or full not in gen._names:
            v = float(defaults.get(slot, 1.)); return (v, v)
        prior = (priors or {}).get(full) or {}
        lim = prior.get('limits')
        if lim is None or lim[0] is None or lim[1] is None or not np.all(np.isfinite(lim)) or lim[0] <= 0.: return None
        return (float(lim[0]), float(lim[1]))

    l0, l1 = limits(q0), limits(q1)
    if l0 is None or l1 is None: return None
    vals = []
    for a in np.linspace(l0[0], l0[1], 9):
        for b in np.linspace(l1[0], l1[1], 9):
            if apmode == 'qparqper': qpar, qper = a, b
            elif apmode == 'qiso': qpar, qper = a, a
            elif apmode == 'qap': qpar, qper = a**(1. - eta), a**(-eta)
            else: qpar, qper = a * b**(1. - eta), a * b**(-eta)
            qap = qpar / qper
            fap = np.sqrt(1. + np.append(mu, [0., 1.])**2 * (1. / qap**2 - 1.))
            vals.append(np.log10(fap / qper))
    vals = np.concatenate(vals)
    return float(vals.min()) - 1e-9, float(vals.max()) + 1e-9


def lower_theory(gen, theory, defaults, names=None, priors=None):
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
        fold = None
        if theory.get('s') is not None:
            # correlation function multipoles (reference full_shape.py:336-365, :553-574 through get_corr, base.py:46-136): the Kaiser tables are
            # evaluated on kin and the (linear) resampling + FFTLog + interpolation onto s is folded into the term matrices at plan time
            from .fftlog import fftlog_operator
            if theory.get('fftlog', 'fold') != 'fold': raise ValueError("Kaiser correlation function multipoles: only theory['fftlog'] = 'fold'")
            interp_order = {'linear': 1, 'cubic': 3}.get(theory.get('interp_order', 1), theory.get('interp_order', 1))   # reference base.py:53-57
            if interp_order not in (1, 3): raise ValueError('interp_order must be one of [1, 3]')
            k = np.geomspace(1e-4, 0.6, int(300. / interp_order + 0.5)) if theory.get('kin') is None else f64(theory['kin'])  # reference base.py:61-67
            s_out = f64(theory['s'])
            T = fftlog_operator(k, ells, s_out, interp_order=interp_order)
            fold = np.zeros((len(ells) * k.size, len(ells) * s_out.size))
            for ill in range(len(ells)): fold[ill * k.size:(ill + 1) * k.size, ill * s_out.size:(ill + 1) * s_out.size] = T[ill]
        else:
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
        from .spline import is_uniform
        uniform = bool(is_uniform(np.log10(k_t)))
        interp = {'scipy': 0, 'not-a-knot': 0, 'interpax': 1, 'local': 1}.get(str(theory.get('interp', tpl.get('interp', 'scipy'))), None)
        if interp is None:
            raise ValueError("theory['interp'] must be 'scipy' (not-a-knot C2 spline, the reference's NumPy branch jax.py:263-265) or 'interpax' "
                             "(C1 local cubic, the reference's jax branch jax.py:258-261)")
        pad = 28 if interp == 0 else 2
        # largest window of template knots a point INSIDE the prior box can reach (the kernel's per-warp shared memory is sized for it): the AP
        # factor log10(F(mu) / qper) is bounded from the limits of the AP parameters; without finite limits the whole grid
        nw_max = nt
        if uniform:
            span = _ap_log_span(apmode, q0, q1, gen, defaults, priors, float(tpl.get('eta', 1. / 3.)), mu)
            if span is not None:
                dx = (np.log10(k_t[-1]) - np.log10(k_t[0])) / (nt - 1)
                nw_max = int(min(nt, np.ceil((np.log10(k.max()) - np.log10(k.min()) + span[1] - span[0]) / dx) + 2 * pad + 6))
        nw_max = max(nw_max, 8)
        gen._kaiser = dict(head=[nk, nell, nmu, nt, APMODES[apmode], (ell0 << 8) | int(uniform), ntab], interp=interp, pad=pad, nw_max=nw_max,
                           base=np.concatenate([[float(tpl.get('eta', 1. / 3.)), a, float(tpl['f_fid']), 0.], k, mu, wmu.ravel(), pk, Tt, Lt,
                                                spline_constants(np.log10(k_t)), [np.log10(k).min(), np.log10(k).max()]]),
                           derived=np.concatenate([np.log10(k), np.append(np.diff(np.log10(k_t)), 0.), _converged_pivots(spline_constants(np.log10(k_t)), nt, uniform)]),
                           ntab=ntab, nin=nin, fold=fold)
        # default output segments = the four plain tables; Lowered may replace them by merged combinations (kaiser_set_segments)
        kaiser_set_segments(gen, [dict(name='pk_dd', flag=0, combs=[(0, 1., ())]), dict(name='pk_dt', flag=0, combs=[(1, 1., ())]),
                                  dict(name='pk_tt', flag=0, combs=[(2, 1., ())]), dict(name='pk11_l0', flag=1, combs=[(0, 1., ())])])
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
            interp_order = {'linear': 1, 'cubic': 3}.get(theory.get('interp_order', 1), theory.get('interp_order', 1))   # reference base.py:53-56
            if interp_order not in (1, 3): raise ValueError('interp_order must be one of [1, 3]')
            if interp_order == 3 and theory.get('fftlog', 'fold') != 'fold':
                raise ValueError("interp_order=3 of the P -> xi resampling needs theory['fftlog'] = 'fold' (the explicit FFTLog kernel resamples linearly)")
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
            T = fftlog_operator(k, ells, s_out, interp_order=interp_order)
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
