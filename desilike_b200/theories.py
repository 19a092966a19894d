"""Host-side mirrors of the reference's theory calculators on the batched path.  They are *declarative*: they hold the
constants and the parameter collection (names, defaults and priors copied from the reference's YAML) and describe
themselves as the ``theory`` / ``terms`` parts of a likelihood spec; all per-point arithmetic happens in the CUDA kernels.

Mirrored (same class names, argument meaning, parameter names):
  ShapeFitPowerSpectrumTemplate, BAOPowerSpectrumTemplate          reference theories/galaxy_clustering/power_template.py:69-172, :340-376, :690-761
  KaiserTracerPowerSpectrumMultipoles, EFTLikeKaiserTracer...       reference theories/galaxy_clustering/full_shape.py:144-207, :466-550, :577-640
  DampedBAOWigglesTracerPowerSpectrumMultipoles / ...CorrelationFunctionMultipoles   reference theories/galaxy_clustering/bao.py:96-151, :424-541, :805-924

The fiducial linear power spectrum comes from cosmoprimo/CLASS in the reference (power_template.py:49-66); here the
caller provides it (callable or tabulated) -- it is a constant of the path, not part of its arithmetic.
"""
import re

import numpy as np

from .parameter import Parameter, ParameterCollection

_Q = dict(prior=dict(limits=[0.8, 1.2]), ref=dict(limits=[0.99, 1.01]), value=1., delta=0.008)
_AP_PARAMS = {'qparqper': ['qpar', 'qper'], 'qiso': ['qiso'], 'qap': ['qap'], 'qisoqap': ['qiso', 'qap']}


class BaseCalculator(object):
    """Common parameter plumbing (subset of reference desilike/base.py:1119-1322)."""

    @property
    def all_params(self):
        return self.params

    @property
    def varied_params(self):
        return self.all_params.select(varied=True)


def _tabulate(fun, k):
    if callable(fun): return np.asarray(fun(k), dtype='f8')
    fun = np.asarray(fun, dtype='f8')
    if fun.shape != np.shape(k): raise ValueError('tabulated fiducial spectrum must match the template k-grid')
    return fun


class BasePowerSpectrumTemplate(BaseCalculator):

    def __init__(self, pk_dd_fid, f_fid, pknow_dd_fid=None, k=None, z=1., apmode='qparqper', eta=1. / 3., only_now=False, rs_drag=None):
        self.pk_dd_fid, self.pknow_dd_fid, self.f_fid = pk_dd_fid, pknow_dd_fid, float(f_fid)
        self.k = None if k is None else np.asarray(k, dtype='f8')
        self.z, self.apmode, self.eta, self.only_now, self.rs_drag_fid = z, apmode, float(eta), bool(only_now), rs_drag
        if apmode not in _AP_PARAMS: raise ValueError('unknown apmode {}'.format(apmode))
        self.params = ParameterCollection([Parameter(name, **_Q) for name in _AP_PARAMS[apmode]])

    def template_spec(self, k):
        """Constants on the template grid ``k`` (chosen by the PT theory, reference full_shape.py:29)."""
        if self.k is not None: k = self.k
        spec = dict(k=k, pk_dd_fid=_tabulate(self.pk_dd_fid, k), f_fid=self.f_fid, apmode=self.apmode, eta=self.eta, only_now=self.only_now)
        spec['pknow_dd_fid'] = _tabulate(self.pknow_dd_fid, k) if self.pknow_dd_fid is not None else spec['pk_dd_fid']
        return spec


class ShapeFitPowerSpectrumTemplate(BasePowerSpectrumTemplate):
    """pk_dd = pk_dd_fid * exp(dm / a * tanh(a ln(k / kp)) + dn ln(k / kp)),  f = f_fid * df  (reference power_template.py:747-761)."""

    def __init__(self, *args, kp=0.03, a=0.6, **kwargs):
        super().__init__(*args, **kwargs)
        self.kp, self.a = float(kp), float(a)
        for p in [Parameter('dm', value=0., prior=dict(limits=[-3., 3.]), ref=dict(limits=[-0.01, 0.01]), delta=0.01),
                  Parameter('dn', value=0., fixed=True, prior=dict(limits=[-0.5, 0.5])),
                  Parameter('df', value=1., prior=dict(limits=[0., 2.]), ref=dict(limits=[0.95, 1.05]), delta=0.02)]:
            self.params.set(p)

    def template_spec(self, k):
        return dict(super().template_spec(k), kind='shapefit', a=self.a, kp=self.kp)


class BAOPowerSpectrumTemplate(BasePowerSpectrumTemplate):
    """Fixed linear template, f = f_fid * df (reference power_template.py:372-376)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params.set(Parameter('df', value=1., fixed=True, prior=dict(limits=[0., 2.])))

    def template_spec(self, k):
        return dict(super().template_spec(k), kind='bao')


# ---------------------------------------------------------------------------------------------------------------------
class BaseTracerPowerSpectrumMultipoles(BaseCalculator):
    _klim = (1e-3, 1., 500)   # reference full_shape.py:19
    is_correlation = False

    def __init__(self, template=None, pt=None, k=None, ells=(0, 2, 4), mu=8, shotnoise=1e4, tracer=None):
        self.template, self.pt = template, pt
        self.k = np.linspace(0.01, 0.2, 101) if k is None else np.asarray(k, dtype='f8')   # reference base.py:16-19
        self.ells, self.mu, self.nd, self.tracer = tuple(ells), int(mu), 1. / float(shotnoise), tracer
        self.params = ParameterCollection()
        self._pt = None   # the perturbation-theory parameters are created ONCE (updates through all_params must stick)

    def _ns(self, name):
        return '{}.{}'.format(self.tracer, name) if self.tracer else name

    def set_k(self, k, ells):
        """Called by the window (reference window.py:443: ``theory.init.update(k=kin, ells=ellsin)``)."""
        self.k, self.ells = np.asarray(k, dtype='f8'), tuple(ells)

    @property
    def all_params(self):
        if self.pt is None and self._pt is None: self._pt = self._pt_params()
        base = self.pt.all_params if self.pt is not None else (self.template.params + self._pt)
        return base + self.params

    def _pt_params(self):
        return ParameterCollection()

    @property
    def nin(self):
        return len(self.ells) * len(self.k)


class KaiserTracerPowerSpectrumMultipoles(BaseTracerPowerSpectrumMultipoles):
    """P_ell = b1^2 pk_dd + 2 b1 pk_dt + pk_tt + delta_{ell 0} sn0 / nd  (reference full_shape.py:545-550; tables :488-500)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params = ParameterCollection([
            Parameter('b1', prior=dict(limits=[0., 4.]), ref=dict(limits=[1., 2.]), namespace=self.tracer),
            Parameter('sn0', prior=dict(dist='norm', loc=0., scale=1000.), ref=dict(dist='norm', loc=0., scale=0.1), namespace=self.tracer)])

    def _pt_params(self):
        return ParameterCollection([Parameter('sigmapar', value=0., fixed=True, prior=dict(limits=[0., 10.])),
                                    Parameter('sigmaper', value=0., fixed=True, prior=dict(limits=[0., 10.]))])

    def theory_spec(self):
        if self.pt is not None:
            return self.pt.theory_spec(self)
        k_t = np.geomspace(min(self._klim[0], self.k[0] / 2), max(self._klim[1], self.k[-1] * 2), self._klim[2])  # reference full_shape.py:29
        names = self.all_params.names()
        params = [n for n in names if n.split('.')[-1] in ('qpar', 'qper', 'qiso', 'qap', 'dm', 'dn', 'df', 'sigmapar', 'sigmaper')]
        return dict(kind='kaiser', k=self.k, ells=self.ells, nmu=self.mu, template=self.template.template_spec(k_t), params=params,
                    defaults={p.basename: p.value for p in self.all_params})

    def terms(self):
        nk = len(self.k)
        e0 = np.concatenate([np.full(nk, float(ell == 0)) for ell in self.ells])
        b1, sn0 = self._ns('b1'), self._ns('sn0')
        return [dict(scale=1., factors=[(b1, 2)], table='pk_dd'), dict(scale=2., factors=[(b1, 1)], table='pk_dt'),
                dict(scale=1., factors=[], table='pk_tt'), dict(scale=1. / self.nd, factors=[(sn0, 1)], table=None, colvec=e0)]


def _xi_fold(kin, ells, s, interp_order=1):
    """[nell * nkin, nell * ns]: the linear P_ell(kin) -> xi_ell(s) operator of get_corr (reference base.py:127-136), block diagonal in ell."""
    from .fftlog import fftlog_operator
    kin, s = np.asarray(kin, dtype='f8'), np.asarray(s, dtype='f8')
    T = fftlog_operator(kin, ells, s, interp_order=interp_order)
    fold = np.zeros((len(ells) * kin.size, len(ells) * s.size))
    for ill in range(len(ells)): fold[ill * kin.size:(ill + 1) * kin.size, ill * s.size:(ill + 1) * s.size] = T[ill]
    return fold


def _fold_emulated_tables(spec, fold, weighted=None, keep=None):
    """Emulated tables on (kin, ells) -> the same tables behind the P -> xi operator ``fold``, folded into the emulator's output maps (Taylor: coefficient
    matrices; MLP: linear post-map).  ``weighted``: {new name: (source table, weights on P(kin))}; ``keep``: {new name: source table} (default: all)."""
    tables, weighted = {}, weighted or {}
    keep = {name: name for name in spec['tables']} if keep is None else keep
    if spec['kind'] == 'taylor':
        for name, src in keep.items(): tables[name] = np.asarray(spec['tables'][src], dtype='f8') @ fold
        for name, (src, w) in weighted.items(): tables[name] = (np.asarray(spec['tables'][src], dtype='f8') * w) @ fold
    else:
        def post(sl, w=None):   # MLP tables: output indices, optionally with a linear post-map [len(indices), nell * nkin]
            sl, m = sl if isinstance(sl, tuple) else (sl, None)
            if m is None: return (sl, fold if w is None else w[:, None] * fold)
            m = np.asarray(m, dtype='f8')
            return (sl, (m if w is None else m * w) @ fold)
        for name, src in keep.items(): tables[name] = post(spec['tables'][src])
        for name, (src, w) in weighted.items(): tables[name] = post(spec['tables'][src], w)
    return dict(spec, tables=tables)


class KaiserTracerCorrelationFunctionMultipoles(KaiserTracerPowerSpectrumMultipoles):
    """xi_ell(s) = Hankel transforms of the Kaiser tracer P_ell(k) = b1^2 pk_dd + 2 b1 pk_dt + pk_tt evaluated on ``kin`` (reference
    full_shape.py:336-365, :553-574; get_corr base.py:46-136: resampling onto the FFTLog grid -- ``interp_order`` 1 linear / 3 cubic in log10 k --,
    damped high-k extrapolation, FFTLog, interpolation onto s).  No stochastic parameter (reference :573)."""
    is_correlation = True

    def __init__(self, template=None, pt=None, s=None, ells=(0, 2, 4), mu=8, interp_order=1, tracer=None):
        self.interp_order = {'linear': 1, 'cubic': 3}.get(interp_order, interp_order)
        if self.interp_order not in [1, 3]: raise ValueError('interp_order must be one of {}'.format([1, 3]))
        self.kin = np.geomspace(1e-4, 0.6, int(300. / self.interp_order + 0.5))   # reference base.py:61-67
        # pt: an EmulatedCalculator of the Kaiser perturbation-theory tables ON kin (what the reference emulates as ``theory.pt``, e.g.
        # ``Emulator.fit_taylor(KaiserTracerPowerSpectrumMultipoles(template, k=xi.kin, ells=ells))``)
        super().__init__(template=template, pt=pt, k=self.kin, ells=ells, mu=mu, tracer=tracer)
        self.s = np.linspace(20., 200, 101) if s is None else np.asarray(s, dtype='f8')
        self.params = ParameterCollection([Parameter('b1', prior=dict(limits=[0., 4.]), ref=dict(limits=[1., 2.]), namespace=self.tracer)])

    def _weighted_tables(self):
        """Tables entering with weights on P(kin): {new table name: (source table, weights [nell * nkin])} (counter terms of the EFT-like class)."""
        return {}

    def set_k(self, s, ells):
        self.s, self.ells = np.asarray(s, dtype='f8'), tuple(ells)

    @property
    def nin(self):
        return len(self.ells) * len(self.s)

    def theory_spec(self):
        if self.pt is None:
            return dict(super().theory_spec(), s=self.s, kin=self.kin, interp_order=self.interp_order)
        # emulated tables on (kin, ells): fold the P -> xi operator (and the weights of weighted tables) into the emulator's output maps
        return _fold_emulated_tables(self.pt.theory_spec(self), _xi_fold(self.kin, self.ells, self.s, self.interp_order), weighted=self._weighted_tables())

    def terms(self):
        b1 = self._ns('b1')
        return [dict(scale=1., factors=[(b1, 2)], table='pk_dd'), dict(scale=2., factors=[(b1, 1)], table='pk_dt'), dict(scale=1., factors=[], table='pk_tt')]


class EFTLikeKaiserTracerPowerSpectrumMultipoles(KaiserTracerPowerSpectrumMultipoles):
    """Kaiser + EFT-like counter terms ct{ell}_{pow} * (k / kp)^pow * pk11_{ell=0} and stochastic terms sn{ell}_{pow} * (k / kp)^pow / nd
    (reference full_shape.py:577-640), exactly marginalisable (set ``derived='.marg'``)."""
    kp = 1.

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        b1, sn0 = self.params['b1'], self.params['sn0']
        params = [b1]   # order of the reference YAML (full_shape.yaml:112-152): b1, ct*, sn0, sn*_2
        for ell in (0, 2, 4):
            params.append(Parameter('ct{:d}_2'.format(ell), value=0., prior=dict(dist='norm', loc=0., scale=100.), namespace=self.tracer))
        params.append(sn0)
        for ell in (0, 2, 4):
            params.append(Parameter('sn{:d}_2'.format(ell), value=0., prior=dict(dist='norm', loc=0., scale=1000.), namespace=self.tracer))
        self.params = ParameterCollection(params)

    def terms(self):
        terms = super().terms()
        nk = len(self.k)
        for p in self.params:
            match = re.match(r'(ct|sn)(\d+)_(-?\d+)$', p.basename)
            if not match: continue
            base, ell, power = match.group(1), int(match.group(2)), int(match.group(3))
            if ell not in self.ells: continue
            colvec = np.zeros((len(self.ells), nk))
            colvec[self.ells.index(ell)] = (self.k / self.kp)**power
            if base == 'ct':
                terms.append(dict(scale=1., factors=[(p.name, 1)], table='pk11_l0', colvec=colvec.ravel()))
            else:
                terms.append(dict(scale=1. / self.nd, factors=[(p.name, 1)], table=None, colvec=colvec.ravel()))
        # reference behaviour: 'sn0' is added by KaiserTracer.calculate (full_shape.py:549) AND by the stochastic matrix (:634)
        terms[3] = dict(terms[3], scale=2. / self.nd)
        return terms


class EFTLikeKaiserTracerCorrelationFunctionMultipoles(KaiserTracerCorrelationFunctionMultipoles):
    """Hankel transforms of the EFT-like Kaiser tracer P_ell(k): Kaiser + counter terms ct{ell}_{pow} (k / kp)^pow pk11_{ell=0} on ``kin`` (reference
    full_shape.py:577-634, :664-684; no stochastic parameters, :683).  The counter terms stay linear after the P -> xi operator, so they can be
    analytically marginalised (``derived='.marg'``): their templates are feature-space weights applied before the folded operator (``featvec``)."""
    kp = 1.

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for ell in (0, 2, 4):
            self.params.set(Parameter('ct{:d}_2'.format(ell), value=0., prior=dict(dist='norm', loc=0., scale=100.), namespace=self.tracer))

    def _counter_terms(self):
        """[(parameter name, weights on pk11_{ell=0}(kin) [nell * nkin])] (reference full_shape.py:588-618: (k / kp)^pow on the ell block)."""
        out, nk = [], len(self.kin)
        for p in self.params:
            match = re.match(r'ct(\d+)_(-?\d+)$', p.basename)
            if not match: continue
            ell, power = int(match.group(1)), int(match.group(2))
            if ell not in self.ells: continue
            featvec = np.zeros((len(self.ells), nk))
            featvec[self.ells.index(ell)] = (self.kin / self.kp)**power
            out.append((p.name, featvec.ravel()))
        return out

    def _weighted_tables(self):
        return {'pk11_l0@' + name: ('pk11_l0', w) for name, w in self._counter_terms()}

    def terms(self):
        terms = super().terms()
        for name, w in self._counter_terms():
            if self.pt is None: terms.append(dict(scale=1., factors=[(name, 1)], table='pk11_l0', featvec=w))
            else: terms.append(dict(scale=1., factors=[(name, 1)], table='pk11_l0@' + name))   # weights folded into the emulator's output map
        return terms


class TableBiasTracerPowerSpectrumMultipoles(BaseTracerPowerSpectrumMultipoles):
    """Generic table model ``P_ell(k) = interp1d(k, pt.k, sum_i table[ell, :, i] * m_i(bias))`` behind a Taylor emulator of ``table`` [nell, nk_pt, nm]:
    the structure shared by the reference's perturbation-theory tracers (``combine_bias_terms_poles``, reference full_shape.py:1168, :1214, :1479,
    :1711, :1976 ...).  ``monomials``: one tuple of parameter basenames per table slice (``('b1', 'b1')`` = b1^2, ``()`` = 1), optionally
    ``(scale, names)``; ``params``: the bias ``Parameter`` objects.  Parameters entering linearly can be tagged ``derived='.marg'``."""
    table_name = 'pktable'

    def __init__(self, pt, monomials=None, params=(), k=None, ells=(0, 2, 4), shotnoise=1e4, tracer=None, table=None):
        super().__init__(template=None, pt=pt, k=k, ells=ells, shotnoise=shotnoise, tracer=tracer)
        if table is not None: self.table_name = str(table)
        if monomials is not None: self._monomials = [tuple(m) for m in monomials]
        self.params = ParameterCollection([Parameter(p.basename, value=p.value, fixed=p.fixed, prior=p.prior, ref=p.ref, derived=p.derived, namespace=self.tracer) for p in params])

    def _monomial(self, i):
        """-> (scale, names) of table slice i."""
        mono = self._monomials[i]
        if len(mono) == 2 and not isinstance(mono[0], str): return float(mono[0]), tuple(mono[1])
        return 1., tuple(mono)

    def theory_spec(self):
        from scipy import interpolate
        spec, deriv = self.pt.table_spec(self.table_name)   # Taylor: derivatives [T, nell_pt, nk_pt, nm]; MLP: the output's yshape
        kpt, ells_pt = self.pt.k, self.pt.ells
        yshape = tuple(deriv.shape[1:]) if spec['kind'] == 'taylor' else tuple(deriv)
        if yshape != (len(ells_pt), len(kpt), len(self._monomials)):
            raise ValueError('{} emulator must have yshape (nell, nk, {:d}), got {}'.format(self.table_name, len(self._monomials), yshape))
        if self.k[0] < kpt[0] or self.k[-1] > kpt[-1]: raise ValueError('requested k outside the emulated pt.k range')
        index = [ells_pt.index(ell) for ell in self.ells]
        # interp1d(k, pt.k, .) (reference jax.py:211-265; SciPy not-a-knot cubic in the NumPy fallback) as a matrix [nk, nk_pt]
        C = interpolate.interp1d(kpt, np.eye(len(kpt)), kind='cubic', axis=0, fill_value='extrapolate')(self.k)
        for i in range(len(self._monomials)):
            name = '{}_{:d}'.format(self.table_name, i)
            if spec['kind'] == 'taylor':
                spec['tables'][name] = np.einsum('tlk,qk->tlq', deriv[..., i][:, index], C).reshape(deriv.shape[0], -1)
            else:   # indices of [ell, :, i] in the flat network output, then the block-diagonal interpolation over the multipoles
                idx = np.arange(int(np.prod(yshape))).reshape(yshape)[index, :, i].ravel()
                post = np.zeros((len(index) * len(kpt), len(index) * len(self.k)))
                for l in range(len(index)): post[l * len(kpt):(l + 1) * len(kpt), l * len(self.k):(l + 1) * len(self.k)] = C.T
                spec['tables'][name] = (idx, post)
        return spec

    def terms(self):
        terms = []
        for i in range(len(self._monomials)):
            scale, names = self._monomial(i)
            factors = {}
            for name in names: factors[self._ns(name)] = factors.get(self._ns(name), 0) + 1
            terms.append(dict(scale=scale, factors=list(factors.items()), table='{}_{:d}'.format(self.table_name, i)))
        return terms


class LPTVelocileptorsTracerPowerSpectrumMultipoles(TableBiasTracerPowerSpectrumMultipoles):
    """Velocileptors-style table model behind an emulator file (reference full_shape.py:1182-1186, :1252-1313, standard bias basis):

        P_ell(k) = interp1d(k, pt.k, sum_i pktable[ell, :, i] * m_i),
        m = [1, b1, b1^2, b2, b1 b2, b2^2, bs, b1 bs, b2 bs, bs^2, b3, b1 b3, alpha0, alpha2, alpha4, alpha6, sn0 / nd, sn2 / nd, sn4 / nd]

    with ``pktable`` [nell, nk_pt, 19] predicted by a Taylor emulator (``pt``: ``EmulatedCalculator`` with an engine 'pktable'; velocileptors
    itself is third party and absent).  The cubic (not-a-knot) re-interpolation pt.k -> k is linear in the ordinates, so it is folded
    into the 19 table maps at plan time; per point the CUDA path evaluates 19 bias monomials x table contractions.  alpha* and sn* enter
    linearly: tag them ``derived='.marg'`` for analytic marginalisation (each becomes a theta-dependent Jacobian row of the plan)."""
    _monomials = [(), ('b1',), ('b1', 'b1'), ('b2',), ('b1', 'b2'), ('b2', 'b2'), ('bs',), ('b1', 'bs'), ('b2', 'bs'), ('bs', 'bs'), ('b3',), ('b1', 'b3'),
                  ('alpha0',), ('alpha2',), ('alpha4',), ('alpha6',), ('sn0',), ('sn2',), ('sn4',)]

    # reference full_shape.py:1075-1090 (get_physical_stochastic_settings)
    _physical_settings = {'BGS': dict(fsat=0.15, sigv=150 * 10**(1 / 3) * (1 + 0.2)**0.5 / 70.), 'LRG': dict(fsat=0.15, sigv=150 * 10**(1 / 3) * (1 + 0.8)**0.5 / 70.),
                          'ELG': dict(fsat=0.10, sigv=150 * 2.1**0.5 / 70.), 'QSO': dict(fsat=0.03, sigv=150 * 10**(0.7 / 3) * 2.4**0.5 / 70.), None: dict(fsat=0.1, sigv=5.)}

    def __init__(self, pt, k=None, ells=(0, 2, 4), shotnoise=1e4, tracer=None, prior_basis='standard', tracer_type=None, fsat=None, sigv=None):
        """``prior_basis='physical'`` (the reference's default, full_shape.py:1275, :1299-1307): the sampled parameters are b1p .. sn4p and
            b1 = b1p / sigma8 - 1,  b2 = b2p / sigma8^2,  bs = bsp / sigma8^2,  b3 = b3p / sigma8^3,     f = fsigma8 / sigma8,
            alpha0 = (1 + b1)^2 alpha0p,  alpha2 = f (1 + b1) (alpha0p + alpha2p),  alpha4 = f (f alpha2p + (1 + b1) alpha4p),  alpha6 = f^2 alpha4p,
            sn_i = sn_ip snd (fsat if i > 0) sigv^i
        with sigma8 / fsigma8 taken from the emulator (scalar Taylor engines, or fixed entries).  On the CUDA path b1 .. b3, sigma8 and f become
        derived-expression parameters (evaluated per point by the derived-parameter kernel) and the alpha / sn slices expand into monomial terms
        in (b1, f) that stay linear in alpha*p / sn*p: these can still be tagged ``derived='.marg'``."""
        if prior_basis not in (None, 'standard', 'physical'): raise ValueError("prior_basis must be 'standard' or 'physical'")
        BaseTracerPowerSpectrumMultipoles.__init__(self, template=None, pt=pt, k=k, ells=ells, shotnoise=shotnoise, tracer=tracer)
        self.is_physical_prior = prior_basis == 'physical'
        self.nd = 1e-4                      # reference full_shape.py:1154: the shot-noise terms are normalised by 1e-4 whatever `shotnoise`
        self.fsat = self.snd = self.sigv = 1.
        norm = lambda scale: dict(dist='norm', loc=0., scale=scale)
        P = lambda name, **kw: Parameter(name, namespace=self.tracer, **kw)
        if not self.is_physical_prior:
            # reference full_shape.yaml:476-578
            self.params = ParameterCollection([
                P('b1', value=1., prior=dict(limits=[-1., 10.]), ref=dict(limits=[0.4, 0.6])), P('b2', value=0., prior=norm(10.), ref=norm(0.5)),
                P('bs', value=0., prior=norm(5.), ref=norm(0.5)), P('b3', value=0., fixed=True, prior=norm(5.), ref=norm(0.5)),
                P('alpha0', value=0., prior=norm(30.), ref=norm(1.)), P('alpha2', value=0., prior=norm(50.), ref=norm(1.)), P('alpha4', value=0., prior=norm(50.), ref=norm(1.)),
                P('alpha6', value=0., fixed=True, prior=norm(50.), ref=norm(1.)),
                P('sn0', value=0., prior=norm(4.), ref=norm(0.1)), P('sn2', value=0., prior=norm(100.), ref=norm(0.1)), P('sn4', value=0., prior=norm(500.), ref=norm(0.1))])
            return
        settings = dict(self._physical_settings[str(tracer_type).upper() if tracer_type is not None else None])
        self.fsat = settings['fsat'] if fsat is None else float(fsat)
        self.sigv = settings['sigv'] if sigv is None else float(sigv)
        self.snd = float(shotnoise) * self.nd
        ns = self._ns
        s8 = '(' + pt.scalar_expression('sigma8') + ')'
        fs8 = '(' + pt.scalar_expression('fsigma8') + ')'
        ref1 = dict(dist='norm', loc=0., scale=1.)
        # reference full_shape.py:1114-1131 (priors of the physical basis) + the derived standard-basis parameters
        self.params = ParameterCollection([
            P('b1p', value=1., prior=dict(dist='uniform', limits=[0., 3.]), ref=dict(dist='norm', loc=1., scale=0.1)),
            P('b2p', value=0., prior=norm(5.), ref=ref1), P('bsp', value=0., prior=norm(5.), ref=ref1), P('b3p', value=0., fixed=True, prior=norm(5.), ref=ref1),
            P('alpha0p', value=0., prior=norm(12.5), ref=ref1), P('alpha2p', value=0., prior=norm(12.5), ref=ref1), P('alpha4p', value=0., prior=norm(12.5), ref=ref1),
            P('sn0p', value=0., prior=norm(2.), ref=ref1), P('sn2p', value=0., prior=norm(5.), ref=ref1), P('sn4p', value=0., prior=norm(5.), ref=ref1),
            P('sigma8', derived=s8), P('f', derived=fs8 + ' / {' + ns('sigma8') + '}'),
            P('b1', derived='{' + ns('b1p') + '} / {' + ns('sigma8') + '} - 1.'), P('b2', derived='{' + ns('b2p') + '} / {' + ns('sigma8') + '}**2'),
            P('bs', derived='{' + ns('bsp') + '} / {' + ns('sigma8') + '}**2'), P('b3', derived='{' + ns('b3p') + '} / {' + ns('sigma8') + '}**3')])

    def _monomial(self, i):
        mono = self._monomials[i]
        return (1. / self.nd if mono and mono[0].startswith('sn') else 1.), mono

    def terms(self):
        if not self.is_physical_prior: return super().terms()
        ns, tab = self._ns, lambda i: '{}_{:d}'.format(self.table_name, i)
        terms = []
        for i in range(12):   # deterministic bias monomials in the derived (b1, b2, bs, b3)
            scale, names = self._monomial(i)
            factors = {}
            for name in names: factors[ns(name)] = factors.get(ns(name), 0) + 1
            terms.append(dict(scale=scale, factors=list(factors.items()), table=tab(i)))
        b1, f = ns('b1'), ns('f')
        T = lambda i, scale, *factors: terms.append(dict(scale=scale, factors=[(n, p) for n, p in factors], table=tab(i)))
        a0, a2, a4 = ns('alpha0p'), ns('alpha2p'), ns('alpha4p')
        T(12, 1., (a0, 1)); T(12, 2., (b1, 1), (a0, 1)); T(12, 1., (b1, 2), (a0, 1))                       # alpha0 = (1 + b1)^2 alpha0p
        for a in (a0, a2): T(13, 1., (f, 1), (a, 1)); T(13, 1., (f, 1), (b1, 1), (a, 1))                   # alpha2 = f (1 + b1) (alpha0p + alpha2p)
        T(14, 1., (f, 2), (a2, 1)); T(14, 1., (f, 1), (a4, 1)); T(14, 1., (f, 1), (b1, 1), (a4, 1))      # alpha4 = f (f alpha2p + (1 + b1) alpha4p)
        T(15, 1., (f, 2), (a4, 1))                                                                         # alpha6 = f^2 alpha4p
        for j, i in enumerate([0, 2, 4]):                                                                  # sn_i / nd = sn_ip snd (fsat) sigv^i / nd
            T(16 + j, self.snd * (self.fsat if i > 0 else 1.) * self.sigv**i / self.nd, (ns('sn{:d}p'.format(i)), 1))
        return terms


# ---------------------------------------------------------------------------------------------------------------------
class _BaseBAOTracer(BaseCalculator):

    def __init__(self, template=None, ells=(0, 2), mu=10, mode='', smoothing_radius=15., tracer=None, pt=None):
        # pt: an ``EmulatedCalculator`` of this theory's multipoles without broadband ('power' / 'corr' engine, e.g. from ``Emulator.fit_taylor``)
        self.pt = pt
        if pt is None and template is None: raise ValueError('provide a template (direct theory) or pt (emulated multipoles)')
        self.template, self.ells, self.mu, self.mode, self.smoothing_radius, self.tracer = template, tuple(ells), int(mu), str(mode), float(smoothing_radius), tracer
        if self.mode not in ('', 'recsym', 'reciso'): raise ValueError('Reconstruction mode {} must be one of ["", "recsym", "reciso"]'.format(mode))
        self.params = ParameterCollection([
            Parameter('b1', prior=dict(limits=[0.2, 4.]), ref=dict(limits=[1.5, 2.5]), namespace=tracer),
            Parameter('dbeta', value=1., prior=dict(limits=[0.7, 1.3]), ref=dict(limits=[0.8, 1.2]), namespace=tracer),
            Parameter('sigmas', value=0., prior=dict(limits=[0., 10.]), ref=dict(limits=[0., 1.]), namespace=tracer),
            Parameter('sigmapar', value=9., fixed=True, prior=dict(limits=[0.1, 10.]), namespace=tracer),
            Parameter('sigmaper', value=6., fixed=True, prior=dict(limits=[0.1, 10.]), namespace=tracer)])

    @property
    def all_params(self):
        if self.pt is not None:   # the emulator carries the template / damping parameters; this class adds the broadband ones
            own = ParameterCollection([p for p in self.params if p.basename.startswith('al')])
            return self.pt.params + own
        return self.template.params + self.params

    def _emulated_spec(self, table, x):
        spec = self.pt.theory_spec()
        if table not in spec['tables']: raise ValueError('the emulator has no {!r} engine (engines: {})'.format(table, sorted(spec['tables'])))
        if len(self.pt.k) != len(x) or not np.allclose(self.pt.k, x) or tuple(self.pt.ells) != tuple(self.ells):
            raise ValueError('the emulator was built on a different grid / multipoles than requested')
        return spec

    def _theory_common(self):
        k_t = np.geomspace(1e-4, 1., 2000) if self.template.k is None else self.template.k   # reference bao.py:67,79
        names = self.all_params.names()
        params = [n for n in names if n.split('.')[-1] in ('qpar', 'qper', 'qiso', 'qap', 'df', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper')]
        return dict(ells=self.ells, nmu=self.mu, template=self.template.template_spec(k_t), mode=self.mode, smoothing_radius=self.smoothing_radius,
                    params=params, defaults={p.basename: p.value for p in self.all_params})

    def _broadband_terms(self, x, pivot, powers):
        terms = []
        for ill, ell in enumerate(self.ells):
            for power in powers:
                name = 'al{:d}_{:d}'.format(ell, power)
                if name not in self.params: continue
                colvec = np.zeros((len(self.ells), len(x)))
                colvec[ill] = (x / pivot)**power
                terms.append(dict(scale=1., factors=[(self.params[name].name, 1)], table=None, colvec=colvec.ravel()))
        return terms


class DampedBAOWigglesTracerPowerSpectrumMultipoles(_BaseBAOTracer):
    """Damped BAO P_ell(k) (Beutler 2017 / Chen 2023 'standard' model, reference bao.py:117-134) + power-law broadband
    al{ell}_{pow} (k / kp)^pow, pow in [-3, 1] (reference bao.py:467-476, :527-534)."""
    is_correlation = False

    def __init__(self, template=None, k=None, kp=None, **kwargs):
        super().__init__(template, **kwargs)
        self.k = np.linspace(0.01, 0.2, 101) if k is None else np.asarray(k, dtype='f8')
        if self.pt is not None and k is None: self.k = np.asarray(self.pt.k, dtype='f8')
        if kp is None:
            if template is None or template.rs_drag_fid is None: raise ValueError('provide kp or the template rs_drag (kp defaults to 2 pi / rs_drag, reference bao.py:493)')
            kp = 2. * np.pi / template.rs_drag_fid
        self.kp = float(kp)
        for ell in self.ells:
            for power in range(-3, 2):
                self.params.set(Parameter('al{:d}_{:d}'.format(ell, power), value=0., ref=dict(limits=[-1e2, 1e2]), delta=0.005, namespace=self.tracer))

    def set_k(self, k, ells):
        self.k, self.ells = np.asarray(k, dtype='f8'), tuple(ells)

    @property
    def nin(self):
        return len(self.ells) * len(self.k)

    def theory_spec(self):
        if self.pt is not None: return self._emulated_spec('power', self.k)
        return dict(self._theory_common(), kind='bao_pk', k=self.k)

    def terms(self):
        return [dict(scale=1., factors=[], table='power')] + self._broadband_terms(self.k, self.kp, range(-3, 2))


class DampedBAOWigglesTracerCorrelationFunctionMultipoles(_BaseBAOTracer):
    """xi_ell(s) from the damped BAO P_ell(k) through FFTLog (reference base.py:46-136) + power-law broadband
    al{ell}_{pow} (s / sp)^pow, pow in [-2, 2], sp = 2 pi / 0.02 (reference bao.py:849-853, :867, :905)."""
    is_correlation = True

    def __init__(self, template=None, s=None, sp=None, fftlog='fold', **kwargs):
        super().__init__(template, **kwargs)
        # fftlog: 'fold' = the (linear) P_ell(kin) -> xi_ell(s) operator is folded into the plan's term matrices; 'kernel' = explicit FFTLog kernel
        self.fftlog = str(fftlog)
        self.s = np.linspace(20., 200., 101) if s is None else np.asarray(s, dtype='f8')
        if self.pt is not None and s is None and self.pt.s is not None: self.s = np.asarray(self.pt.s, dtype='f8')   # ('corr' engine: emulated on s)
        self.sp = 2. * np.pi / 0.02 if sp is None else float(sp)
        for ell in self.ells:
            for power in range(-2, 3):
                self.params.set(Parameter('al{:d}_{:d}'.format(ell, power), value=0., ref=dict(limits=[-1e-3, 1e-3]), delta=0.005, namespace=self.tracer))

    def set_k(self, s, ells):
        self.s, self.ells = np.asarray(s, dtype='f8'), tuple(ells)

    @property
    def nin(self):
        return len(self.ells) * len(self.s)

    def theory_spec(self):
        if self.pt is not None:
            spec = self.pt.theory_spec()
            if 'corr' not in spec['tables'] and 'power' in spec['tables']:
                # the reference's structure: the emulated pt is the P_ell(kin) calculator behind get_corr (bao.py:820-833, base.py:61-67, :127-136)
                if tuple(self.pt.ells) != tuple(self.ells): raise ValueError('the emulator was built on different multipoles than requested')
                return _fold_emulated_tables(spec, _xi_fold(self.pt.k, self.ells, self.s), keep={'corr': 'power'})
            return self._emulated_spec('corr', self.s)
        return dict(self._theory_common(), kind='bao_xi', s=self.s, fftlog=self.fftlog)

    def terms(self):
        return [dict(scale=1., factors=[], table='corr')] + self._broadband_terms(self.s, self.sp, range(-2, 3))
