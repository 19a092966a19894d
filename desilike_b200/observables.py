"""Host-side mirrors of the reference's two-point observables and window on the batched path (declarative: constants only).

  TracerPowerSpectrumMultipolesObservable            reference observables/galaxy_clustering/power_spectrum.py:22-121, :400-404
  TracerCorrelationFunctionMultipolesObservable       reference observables/galaxy_clustering/correlation_function.py:380-381
  window handling (ndarray ``wmatrix`` branch, ``kin`` / ``kinrebin`` / ``kinlim`` rebinning, shot noise in / out, k-mask)
                                                     reference observables/galaxy_clustering/window.py:207-473
File-based inputs (pypower / pycorr / lsstypes containers) are out of scope: arrays only.
"""
import numpy as np


def matrix_lininterp(xin, xout):
    """Matrix for linear interpolation (reference desilike/utils.py:646-657)."""
    toret = np.zeros((len(xin), len(xout)), dtype='f8')
    for iout, xo in enumerate(xout):
        iin = np.searchsorted(xin, xo, side='right') - 1
        if 0 <= iin < len(xin) - 1:
            frac = (xo - xin[iin]) / (xin[iin + 1] - xin[iin])
            toret[iin, iout] = 1. - frac
            toret[iin + 1, iout] = frac
        elif np.isclose(xo, xin[-1]):
            toret[iin, iout] = 1.
    return toret


def _block_diag(mats):
    import scipy.linalg
    return scipy.linalg.block_diag(*mats)


class BaseTwoPointObservable(object):
    _default_step = 0.01
    params = None   # parameters owned by the observable itself (systematic-template amplitudes)

    def _set_systematic_templates(self, templates, x):
        """Systematic templates added to the windowed theory, ``flat += sum_i syst_i * template_i`` (reference observables/galaxy_clustering/
        window.py:470-473, :717-735; ``get_templates`` :1253-1274, parameters ``syst_{i}`` with value 0 :1279-1283).  ``templates``: list / dict of
        arrays over the stacked multipoles, or callables ``template(ell, x)``.  The amplitudes are linear: tag them ``derived='.marg'`` to
        marginalise them analytically (constant Jacobian rows of the plan)."""
        from .parameter import Parameter, ParameterCollection
        self.templates = {}
        if templates is None: templates = {}
        if not isinstance(templates, dict):
            if not isinstance(templates, (list, tuple)): templates = [templates]
            templates = {'syst_{:d}'.format(i): t for i, t in enumerate(templates)}
        size = sum(len(xx) for xx in x)
        for name, template in templates.items():
            if callable(template): template = np.concatenate([template(ell, xx) for ell, xx in zip(self.ells, x)])
            template = np.ravel(np.asarray(template, dtype='f8'))
            if template.size != size:
                raise ValueError('provided template is size {:d}, but expected {:d} = sum({})'.format(template.size, size, [len(xx) for xx in x]))
            self.templates[name] = template
        self.params = ParameterCollection([Parameter(name, value=0., ref=dict(limits=[-1e-3, 1e-3]), delta=0.005) for name in self.templates])

    def systematic_terms(self):
        return [dict(scale=1., factors=[(name, 1)], table=None, outvec=template) for name, template in (self.templates or {}).items()]

    def _set_binning(self, x, xlim, ells, step_default):
        """(reference window.py:218-292, simplified to arrays): -> list of per-multipole coordinates."""
        if ells is None: ells = list(xlim) if xlim is not None else (0, 2, 4)
        ells = tuple(ells)
        if x is not None:
            if np.ndim(x[0]) == 0: x = [x] * len(ells)
            x = [np.array(xx, dtype='f8') for xx in x]
            if len(x) != len(ells): raise ValueError("provide as many k's as ells")
        masklim = None
        if xlim is not None:
            xlim = dict(xlim)
            if x is not None:
                newx, newells, masklim = [], [], {}
                for ill, ell in enumerate(ells):
                    xx = x[ill]
                    masklim[ell] = np.zeros(len(xx), dtype='?')
                    if ell not in xlim: continue
                    lo, hi, *step = xlim[ell]
                    mask = (xx >= lo) & (xx <= hi)
                    masklim[ell] = mask
                    if mask.any():
                        newx.append(xx[mask]); newells.append(ell)
                x, ells = newx, tuple(newells)
            else:
                if list(ells) != list(xlim): raise ValueError('incompatible ells = {} and lim = {}'.format(ells, list(xlim)))
                x = []
                for ell in ells:
                    lo, hi, *step = xlim[ell]
                    step = step[0] if step else step_default
                    edges = np.arange(lo, hi + step / 2., step=step)
                    x.append((edges[:-1] + edges[1:]) / 2.)
        if x is None:
            x = [np.arange(0.01, 0.2 + step_default / 2., step_default)] * len(ells)
        return x, ells, masklim


class TracerPowerSpectrumMultipolesObservable(BaseTwoPointObservable):

    def __init__(self, data=None, covariance=None, klim=None, k=None, ells=None, wmatrix=None, kin=None, kinrebin=1, kinlim=None, ellsin=None,
                 theory=None, shotnoise=None, wshotnoise=None, transform=None, systematic_templates=None, fiber_collisions=None, name='spectrum2poles'):
        from .theories import KaiserTracerPowerSpectrumMultipoles
        self.name, self.transform = name, transform
        if transform not in (None, 'cubic'): raise ValueError('transform must be one of [None, "cubic"]')
        self.k, self.ells, masklim = self._set_binning(k, klim, ells, self._default_step)
        self._set_systematic_templates(systematic_templates, self.k)
        if theory is None: raise ValueError('provide a theory (the reference default needs a cosmoprimo fiducial)')
        self.theory = theory
        self.matrix_full, self.kmask, self.offset = None, None, None
        if wmatrix is None:   # reference window.py:296-303
            self.ellsin = tuple(self.ells)
            self.kin = np.unique(np.concatenate(self.k, axis=0))
            if not all(kk.shape == self.kin.shape and np.allclose(kk, self.kin) for kk in self.k):
                kmask = [np.searchsorted(self.kin, kk, side='left') for kk in self.k]
                self.kmask = np.concatenate([self.kin.size * i + km for i, km in enumerate(kmask)], axis=0)
        elif isinstance(wmatrix, np.ndarray):   # reference window.py:309-323
            self.ellsin = tuple(ellsin or self.ells)
            matrix_full = np.array(wmatrix, dtype='f8')
            ksize = sum(len(kk) for kk in self.k)
            if matrix_full.shape[0] != ksize:
                raise ValueError('output "wmatrix" size is {:d}, but got {:d} output "k"'.format(matrix_full.shape[0], ksize))
            kin = np.asarray(kin, dtype='f8').flatten()
            self.kin = kin.copy()
            if kinrebin is not None and kinrebin != 1: self.kin = self.kin[::kinrebin]
            if kinlim is not None: self.kin = self.kin[(self.kin >= kinlim[0]) & (self.kin <= kinlim[-1])]
            rebin = _block_diag([matrix_lininterp(self.kin, kin) for ell in self.ellsin])
            self.matrix_full = matrix_full.dot(rebin.T)
        else:
            raise ValueError('only ndarray window matrices are supported on this path (file readers of pypower / lsstypes are out of scope)')
        if fiber_collisions is not None:   # reference window.py:428-438: kernels on (kin, ellsin), folded into matrix_full / offset
            fc = dict(fiber_collisions)
            with_uncorrelated = fc.pop('with_uncorrelated', True)
            corr, unc = tophat_fiber_collisions_power(self.kin, self.ellsin, **fc)
            block = np.block([[corr[io, ii] for ii in range(corr.shape[1])] for io in range(corr.shape[0])])
            if self.matrix_full is None:
                if self.kmask is not None: raise ValueError('fiber_collisions without a window matrix needs the same k for every multipole')
                if with_uncorrelated: self.offset = unc.ravel()
                self.matrix_full = block
            else:
                if with_uncorrelated: self.offset = self.matrix_full.dot(unc.ravel())
                self.matrix_full = self.matrix_full.dot(block)
        self.theory.set_k(self.kin, self.ellsin)
        # shot noise in / out (reference window.py:445-457)
        shotnoise = 0. if shotnoise is None else float(shotnoise)
        self.shotnoisein = np.array([shotnoise * (ell == 0) for ell in self.ellsin])
        base = np.concatenate([np.full_like(kk, (ell == 0), dtype='f8') for ell, kk in zip(self.ells, self.k)])
        self.shotnoiseout = shotnoise * base
        if wshotnoise is not None:
            self.shotnoisein[...] = 0.
            self.shotnoiseout = shotnoise * (base - np.concatenate(wshotnoise, axis=0) if not np.ndim(wshotnoise[0]) == 0 else base - np.asarray(wshotnoise))
        self.covariance = None if covariance is None else np.array(covariance, dtype='f8')
        self._data_params = dict(data) if isinstance(data, dict) else None
        self.flatdata = None
        if data is not None and self._data_params is None:
            flatdata = np.ravel(np.asarray(data, dtype='f8'))
            if masklim is not None:
                sizes = [len(m) for m in masklim.values()]
                if flatdata.size == sum(sizes):
                    chunks = np.split(flatdata, np.cumsum(sizes)[:-1])
                    flatdata = np.concatenate([chunks[list(masklim).index(ell)][masklim[ell]] for ell in self.ells])
            self.flatdata = flatdata

    @property
    def size(self):
        return sum(len(kk) for kk in self.k)

    def observable_spec(self):
        """Window + constants of this observable as the ``obs`` part of a likelihood spec (everything folded to W, offset)."""
        nin = self.theory.nin
        nkin = len(self.kin)
        W = self.matrix_full
        offset = np.zeros(self.size)
        noise_in = np.repeat(self.shotnoisein, nkin)
        if W is not None:
            offset = W.dot(noise_in)
        else:
            offset = noise_in[self.kmask] if self.kmask is not None else noise_in.copy()
            if self.kmask is not None:
                W = np.eye(nin)[self.kmask]
        if self.offset is not None: offset = offset + self.offset
        offset = offset - self.shotnoiseout
        return dict(name=self.name, nin=nin, wmatrix=W, offset=offset, transform=self.transform)


# Hahn et al. 2016 (arXiv:1609.01714, appendix): mixing polynomials H_{l l'}(r), l > l', ascending powers of r = min(k, q) / max(k, q)
_FC_H = {(2, 0): [-1., 0., 1.], (4, 0): [3. / 4., 0., -5. / 2., 0., 7. / 4.], (4, 2): [0., 0., -1., 0., 1.],
         (6, 0): [-5. / 8., 0., 35. / 8., 0., -63. / 8., 0., 33. / 8.], (6, 2): [0., 0., 7. / 4., 0., -9. / 2., 0., 11. / 4.], (6, 4): [0., 0., 0., 0., -1., 0., 1.]}


def tophat_fiber_collisions_power(k, ells, kin=None, ellsin=None, fs=1., Dfc=0.):
    """Top-hat fiber-collision kernels on power-spectrum multipoles (Hahn et al. 2016; reference window.py:967-1049).  With
    W2D(x) = 2 J1(x) / x, output wavenumbers k, integration nodes q = kin (trapezoid weights w_q), r = min(k, q) / max(k, q):

        kernel_uncorrelated[l, k]        = -(2 l + 1) L_l(0) fs (pi Dfc)^2 / k  W2D(k Dfc)
        kernel_correlated[l, l', k, q]   = delta_ll' Interp(q -> k) - fs Dfc^2 / 2  F_ll'(k, q)  q w_q
        F_ll  = min(q / k, 1) W2D(q Dfc) r^l
        F_ll' = min(q / k, 1) W2D(q Dfc) (2 l + 1) / 2 H_{max(l,l') min(l,l')}(r)   where (l >= l' and k >= q) or (l <= l' and k <= q), else 0

    -> (kernel_correlated [nell, nellin, nk, nq], kernel_uncorrelated [nell, nk]); plan-time constants folded into the window matrix / offset."""
    from scipy import special
    from numpy.polynomial import Polynomial
    from numpy.polynomial.legendre import legval
    k = np.asarray(k, dtype='f8')
    q = k if kin is None else np.asarray(kin, dtype='f8')
    ells = tuple(ells)
    ellsin = ells if ellsin is None else tuple(ellsin)
    w2d = lambda x: 2. * special.j1(x) / x
    leg0 = np.array([legval(0., [0.] * ell + [1.]) for ell in ells])
    uncorrelated = -((2. * np.array(ells) + 1.) * leg0)[:, None] * fs * (np.pi * Dfc)**2 / k * w2d(k * Dfc)
    wq = np.empty_like(q)   # trapezoid weights
    if q.size > 2:
        wq[1:-1] = (q[2:] - q[:-2]) / 2.; wq[0] = (q[1] - q[0]) / 2.; wq[-1] = (q[-1] - q[-2]) / 2.
    elif q.size == 2: wq[:] = (q[1] - q[0]) / 2.
    else: wq[:] = 1.
    K, Q = np.meshgrid(k, q, indexing='ij')
    ratio = np.minimum(K, Q) / np.maximum(K, Q)
    common = np.minimum(Q / K, 1.) * w2d(Q * Dfc)
    interp = matrix_lininterp(q, k).T
    correlated = np.empty((len(ells), len(ellsin), k.size, q.size))
    for io, lo in enumerate(ells):
        for ii, li in enumerate(ellsin):
            if lo == li:
                F = common * ratio**lo
            else:
                F = common * (2. * lo + 1.) / 2. * Polynomial(_FC_H[(max(lo, li), min(lo, li))])(ratio)
                F = np.where((K >= Q) if lo > li else (K <= Q), F, 0.)
            correlated[io, ii] = float(lo == li) * interp - fs * Dfc**2 / 2. * F * q * wq
    return correlated, uncorrelated


def tophat_fiber_collisions_correlation(s, ells, ellsin=None, fs=1., Dfc=0., mu_range_cut=False):
    """Top-hat fiber-collision kernels on correlation-function multipoles (Hahn et al. 2016; reference window.py:1192-1249): pairs with
    transverse separation below ``Dfc``, i.e. |mu| > mu_min(s) = sqrt(1 - (Dfc / s)^2), are lost with probability ``fs``.  With
    I[p] = int_{mu_min <= |mu| <= 1} p(mu) dmu:

        kernel_uncorrelated[l, s] = -(2 l + 1) / 2 fs I[L_l],     kernel_correlated[l, l', s] = delta_ll' - (2 l + 1) / 2 fs I[L_l L_l']

    (``mu_range_cut``: the correlated kernel divided by mu_min where mu_min > 0).  Plan-time constants: the observable feeds them to the
    window as ``matrix_diag`` / ``offset``."""
    from numpy.polynomial import Polynomial
    from numpy.polynomial.legendre import leg2poly
    s = np.asarray(s, dtype='f8')
    ells = tuple(ells)
    ellsin = ells if ellsin is None else tuple(ellsin)
    mu_min = np.sqrt(np.clip(1. - (float(Dfc) / s)**2, 0., None))
    leg = lambda ell: Polynomial(leg2poly([0.] * ell + [1.]))

    def lost(poly):   # integral over mu_min <= |mu| <= 1
        Q = poly.integ()
        return Q(1.) - Q(mu_min) + Q(-mu_min) - Q(-1.)

    uncorrelated = -np.array([(2 * lo + 1.) / 2. * fs * lost(leg(lo)) for lo in ells])
    correlated = np.empty((len(ells), len(ellsin), s.size))
    for io, lo in enumerate(ells):
        for ii, li in enumerate(ellsin):
            tmp = float(lo == li) - (2 * lo + 1.) / 2. * fs * lost(leg(lo) * leg(li))
            if mu_range_cut: tmp = np.where(mu_min > 0., tmp / np.where(mu_min > 0., mu_min, 1.), tmp)
            correlated[io, ii] = tmp
    return correlated, uncorrelated


class TracerCorrelationFunctionMultipolesObservable(BaseTwoPointObservable):
    _default_step = 5.

    def __init__(self, data=None, covariance=None, slim=None, s=None, ells=None, theory=None, transform=None, wmatrix=None, sin=None, ellsin=None,
                 matrix_diag=None, offset=None, fiber_collisions=None, systematic_templates=None, name='correlation2poles'):
        """``wmatrix`` [sum_ell ns_ell, nellin * nsin] (ndarray branch of reference window.py:651-676), ``matrix_diag`` [nell, nellin, nsin]
        (per-separation multipole mixing, e.g. the fiber-collision kernel :690-702) and ``offset`` are the constants of
        ``WindowedCorrelationFunctionMultipoles._apply`` (:717-728); all linear, folded into one matrix at plan time."""
        self.name, self.transform = name, None
        self.s, self.ells, masklim = self._set_binning(s, slim, ells, self._default_step)
        self._set_systematic_templates(systematic_templates, self.s)
        if theory is None: raise ValueError('provide a theory')
        self.theory = theory
        self.smask, self.matrix_full, self.matrix_diag, self.offset = None, None, None, None
        if wmatrix is None:
            self.sin, self.ellsin = np.unique(np.concatenate(self.s, axis=0)), tuple(self.ells if ellsin is None else ellsin)
            if not all(ss.shape == self.sin.shape and np.allclose(ss, self.sin) for ss in self.s):
                smask = [np.searchsorted(self.sin, ss, side='left') for ss in self.s]
                self.smask = np.concatenate([self.sin.size * i + sm for i, sm in enumerate(smask)], axis=0)
        else:
            if sin is None: raise ValueError('provide the input separations sin of wmatrix')
            self.sin, self.ellsin = np.asarray(sin, dtype='f8').ravel(), tuple(self.ells if ellsin is None else ellsin)
            self.matrix_full = np.array(wmatrix, dtype='f8')
            if self.matrix_full.shape != (sum(len(ss) for ss in self.s), len(self.ellsin) * self.sin.size):
                raise ValueError('wmatrix must be of shape ({:d}, {:d})'.format(sum(len(ss) for ss in self.s), len(self.ellsin) * self.sin.size))
        if fiber_collisions is not None:   # reference window.py:686-702 (no matrix_full): matrix_diag = kernel_correlated, offset = kernel_uncorrelated
            if wmatrix is not None or matrix_diag is not None or offset is not None:
                raise ValueError('fiber_collisions cannot be combined with wmatrix / matrix_diag / offset here: fold the kernels yourself')
            fc = dict(fiber_collisions)
            with_uncorrelated = fc.pop('with_uncorrelated', True)
            matrix_diag, unc = tophat_fiber_collisions_correlation(self.sin, self.ells, ellsin=ellsin, **fc)
            if with_uncorrelated: offset = unc.ravel()
        if matrix_diag is not None:
            self.matrix_diag = np.asarray(matrix_diag, dtype='f8')
            nout = len(self.ellsin) if self.matrix_full is not None else len(self.ells)
            if self.matrix_diag.ndim != 3 or self.matrix_diag.shape[0] != nout or self.matrix_diag.shape[2] != self.sin.size:
                raise ValueError('matrix_diag must be of shape (nell, nellin, nsin)')
            self.ellsin = tuple(ellsin) if ellsin is not None else (self.ellsin if self.matrix_diag.shape[1] == len(self.ellsin) else None)
            if self.ellsin is None or len(self.ellsin) != self.matrix_diag.shape[1]: raise ValueError('provide ellsin matching matrix_diag')
        if offset is not None: self.offset = np.ravel(np.asarray(offset, dtype='f8'))
        self.theory.set_k(self.sin, self.ellsin)
        self.covariance = None if covariance is None else np.array(covariance, dtype='f8')
        self._data_params = dict(data) if isinstance(data, dict) else None
        self.flatdata = None if (data is None or self._data_params is not None) else np.ravel(np.asarray(data, dtype='f8'))

    @property
    def size(self):
        return sum(len(ss) for ss in self.s)

    def observable_spec(self):
        nin, ns = self.theory.nin, self.sin.size
        W = None
        if self.matrix_diag is not None:   # out[l, s] = sum_l' matrix_diag[l, l', s] theory[l', s]  (reference window.py:719-720)
            nlo, nli = self.matrix_diag.shape[:2]
            W = np.zeros((nlo * ns, nli * ns))
            for lo in range(nlo):
                for li in range(nli): W[lo * ns:(lo + 1) * ns, li * ns:(li + 1) * ns] = np.diag(self.matrix_diag[lo, li])
        if self.matrix_full is not None: W = self.matrix_full if W is None else self.matrix_full.dot(W)
        offset = None if self.offset is None else self.offset.copy()
        if self.smask is not None:
            W = np.eye(nin)[self.smask] if W is None else W[self.smask]
            if offset is not None: offset = offset[self.smask]
        return dict(name=self.name, nin=nin, wmatrix=W, offset=offset, transform=None)
