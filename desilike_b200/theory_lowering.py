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
        gen.idata = np.array([nk, nell, nmu, nt, APMODES[apmode], (ell0 << 8), ntab, total], dtype=np.int32)
        gen.ddata = np.concatenate([[float(tpl.get('eta', 1. / 3.)), a, float(tpl['f_fid']), 0.], k, mu, wmu.ravel(), pk, Tt, Lt,
                                    spline_constants(np.log10(k_t))])
        gen.nfeat = total
        for t, name in enumerate(['pk_dd', 'pk_dt', 'pk_tt', 'pk11_l0']):
            gen.tables[name] = (2 + t * ntab, nin, None)
        return
    raise ValueError('theory kind {!r} not lowered yet'.format(kind))
