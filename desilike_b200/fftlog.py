"""Plan-time constants of the FFTLog P_ell(k) -> xi_ell(s) transform evaluated by csrc/theory_kernels.cu:fftlog_kernel.

What the reference computes: ``BaseTheoryCorrelationFunctionFromPowerSpectrumMultipoles`` (desilike/theories/galaxy_clustering/
base.py:46-136) builds ``cosmoprimo.PowerToCorrelation(k=logspace(-4, 3, 2048), ell, q=0, lowring=True)`` (:76-77) and in
``get_corr`` (:127-136) resamples P_ell from ``kin`` onto that grid (linear in log10 k, high-k extrapolation with Gaussian
damping), transforms, and linearly interpolates onto the requested ``s``.  cosmoprimo is absent; the algorithm
(Hamilton 2000: power-law bias q' = 1.5 + q, low-ringing shift, zero padding to the next power of two >= 2 N) is restated
here for the constants and validated against the defining integral through the oracle tests.
"""
import numpy as np
from scipy.special import loggamma


def kernel_u(ell, z):
    """Mellin transform of sqrt(2 / pi) j_ell: 2^(z - 3/2) Gamma((ell + z) / 2) / Gamma((3 + ell - z) / 2)."""
    z = np.asarray(z, dtype=np.complex128)
    return np.exp(np.log(2.) * (z - 1.5) + loggamma(0.5 * (ell + z)) - loggamma(0.5 * (3. + ell - z)))


def fftlog_constants(kin, ells, s_out, ng=2048, kmin=1e-4, kmax=1e3, q=0., minfolds=2):
    """-> dict of arrays consumed by the kernel (layout documented in csrc/theory_kernels.cu)."""
    kin = np.asarray(kin, dtype=np.float64)
    s_out = np.asarray(s_out, dtype=np.float64)
    k = np.logspace(np.log10(kmin), np.log10(kmax), ng)
    qq = 1.5 + q
    delta = np.log(k[-1] / k[0]) / (ng - 1)
    npad_size = 2**((ng * minfolds - 1).bit_length())
    npad = npad_size - ng
    pad_in_left, pad_out_left = npad // 2, npad - npad // 2
    # resampling kin -> k (reference base.py:62-74, :127-133)
    mask = k > kin[-1]
    jmap = np.full(ng, -1, dtype=np.int32)
    wgt, lh, dh = np.zeros(ng), np.zeros(ng), np.ones(ng)
    lk_mid, lkin = np.log10(k[~mask]), np.log10(kin)
    j = np.clip(np.searchsorted(lkin, lk_mid, side='right') - 1, 0, kin.size - 2)
    jmap[~mask] = j
    wgt[~mask] = (lk_mid - lkin[j]) / (lkin[j + 1] - lkin[j])
    lh[mask] = np.log10(k[mask] / kin[-1])
    dh[mask] = np.exp(-(k[mask] / kin[-1] - 1.)**2 / (2. * (10.)**2))
    slope_inv = 1. / np.log10(kin[-1] / kin[-2])
    pre = k**(3. - qq) / (2. * np.pi)**1.5
    m = np.arange(0, npad_size // 2 + 1)
    u, post, outj, outw, s_grids = [], [], [], [], []
    for ell in ells:
        lnxy = delta / np.pi * np.angle(kernel_u(ell, qq + 1j * np.pi / delta))
        s = np.exp(lnxy - delta) / k[::-1]
        uu = kernel_u(ell, qq + 2j * np.pi / npad_size / delta * m) * np.exp(-2j * np.pi * lnxy / npad_size / delta * m)
        uu[-1] = uu[-1].real
        u.append(np.stack([uu.real, uu.imag], axis=-1))
        post.append(s**(-qq) * (-1.)**(ell // 2))
        if s_out.min() < s[0] or s_out.max() > s[-1]:
            raise ValueError('requested s outside the FFTLog output grid')
        jj = np.clip(np.searchsorted(s, s_out, side='right') - 1, 0, ng - 2)
        outj.append(jj.astype(np.int32))
        outw.append((s_out - s[jj]) / (s[jj + 1] - s[jj]))
        s_grids.append(s)
    tw = np.exp(-2j * np.pi * np.arange(npad_size // 2) / npad_size)
    # interpolation and extrapolation as ONE 2-point rule  x[i] = c0[i] pin[j0[i]] + c1[i] pin[j0[i] + 1]  (prefactor folded in):
    #   k <= kin[-1]:  (1 - wgt) pin[j] + wgt pin[j + 1]
    #   k >  kin[-1]:  (pin[-1] + (pin[-1] - pin[-2]) slope_inv lh) dh
    j0 = np.where(mask, kin.size - 2, jmap).astype(np.int32)
    c0 = np.where(mask, -slope_inv * lh * dh, 1. - wgt) * pre
    c1 = np.where(mask, (1. + slope_inv * lh) * dh, wgt) * pre
    ddata = np.concatenate([c0, c1, np.array(u).ravel(), np.array(post).ravel(),
                            np.stack([tw.real, tw.imag], axis=-1).ravel(), np.array(outw).ravel()])
    idata = np.concatenate([j0, np.array(outj).ravel()]).astype(np.int32)
    return dict(ddata=ddata, idata=idata, ng=ng, npad_size=npad_size, pad_in_left=pad_in_left, pad_out_left=pad_out_left, k=k, s=np.array(s_grids),
                j0=j0, c0=c0, c1=c1, u=[uu[..., 0] + 1j * uu[..., 1] for uu in u], post=post, outj=outj, outw=outw)


def fftlog_operator(kin, ells, s_out, interp_order=1, **kwargs):
    """The whole P_ell(kin) -> xi_ell(s_out) chain of ``get_corr`` (reference desilike/theories/galaxy_clustering/base.py:127-136:
    log-k resampling + damped high-k extrapolation, FFTLog, linear interpolation onto ``s_out``) is LINEAR in P_ell(kin): returns
    the operators ``T`` [nell, nkin, ns] with ``xi_ell(s_out) = P_ell(kin) @ T[ell]``, obtained at plan time by pushing the unit
    vectors through the transform (NumPy real FFTs of 4096 points, once).  The lowering folds ``T`` into the term matrices together
    with the window and the whitening, so the per-point path is the damped-BAO kernel followed by the contraction; the explicit
    FFTLog kernel (csrc/theory_kernels.cu:fftlog_kernel) remains available with ``theory['fftlog'] = 'kernel'``.

    ``interp_order=3`` (reference base.py:53, :131): the resampling below ``kin[-1]`` is the reference's NumPy-branch ``interp1d(..., method=3)``
    (desilike/jax.py:263-265: SciPy's not-a-knot cubic spline in log10 k with extrapolation) -- still linear in the ordinates, so only the
    resampling rows of the operator change; the high-k extrapolation stays the 2-point rule."""
    if interp_order not in (1, 3): raise ValueError('interp_order must be one of [1, 3]')
    fc = fftlog_constants(kin, ells, s_out, **kwargs)
    nkin, ng, npad_size = np.size(kin), fc['ng'], fc['npad_size']
    X = np.zeros((nkin, npad_size))
    idx = np.arange(ng)
    np.add.at(X, (fc['j0'], fc['pad_in_left'] + idx), fc['c0'])
    np.add.at(X, (fc['j0'] + 1, fc['pad_in_left'] + idx), fc['c1'])
    if interp_order == 3:
        from scipy import interpolate
        kin = np.asarray(kin, dtype=np.float64)
        mid = fc['k'] <= kin[-1]
        R = interpolate.interp1d(np.log10(kin), np.eye(nkin), kind='cubic', fill_value='extrapolate', axis=0)(np.log10(fc['k'][mid]))   # [nmid, nkin]
        pre = fc['k']**(3. - (1.5 + kwargs.get('q', 0.))) / (2. * np.pi)**1.5
        X[:, fc['pad_in_left'] + idx[mid]] = (R * pre[mid, None]).T
    spec = np.fft.rfft(X, axis=-1)
    T = []
    for ill in range(len(ells)):
        xi = np.fft.hfft(spec * fc['u'][ill], n=npad_size, axis=-1)[:, fc['pad_out_left']:fc['pad_out_left'] + ng] / npad_size * fc['post'][ill]
        jj, w = fc['outj'][ill], fc['outw'][ill]
        T.append(xi[:, jj] * (1. - w) + xi[:, jj + 1] * w)
    return np.array(T)
