"""CPU restatement of FFTLog ``PowerToCorrelation`` (TEST INFRASTRUCTURE — checker only, never shipped).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

The reference calls ``cosmoprimo.PowerToCorrelation(k, ell=ells, q=0, lowring=True)`` at
reference desilike/theories/galaxy_clustering/base.py:76-77 and applies it in ``get_corr`` (:127-136).
cosmoprimo is a third-party git dependency (pyproject.toml:21, *unpinned*, absent from /root/reference and
from this image), so the algorithm is restated here from its published definition: Hamilton (2000) FFTLog
with power-law bias q' = 1.5 + q, low-ringing shift, zero padding to the next power of two >= 2 N.

PARITY UNPINNED against the real cosmoprimo output (no golden vectors exist in the reference);
pinned *mathematically* against the defining integral restated at reference
desilike/theories/galaxy_clustering/base.py:163-168:
    xi_ell(s) = (-1)^(ell/2) / (2 pi^2) * int k^2 dk P_ell(k) j_ell(k s)
(see tests/test_oracle_math.py: analytic Gaussian pair and brute-force quadrature).
"""
import numpy as np
from scipy.special import loggamma


def spherical_bessel_kernel(ell, z):
    """U_ell(z) = 2^(z - 3/2) Gamma((ell + z) / 2) / Gamma((3 + ell - z) / 2): Mellin transform of sqrt(2/pi) j_ell."""
    z = np.asarray(z, dtype='c16')
    return np.exp(np.log(2.) * (z - 1.5) + loggamma(0.5 * (ell + z)) - loggamma(0.5 * (3. + ell - z)))


class PowerToCorrelation(object):
    """
    ``s, xi = PowerToCorrelation(k, ell=(0, 2), q=0, lowring=True)(pk)`` with ``pk`` of shape (nell, nk);
    returns ``s`` (nell, nk) (one output grid per multipole: the low-ringing shift depends on ell) and ``xi`` (nell, nk).
    """

    def __init__(self, k, ell=0, q=0., lowring=True, minfolds=2):
        self.k = np.asarray(k, dtype='f8')
        self.scalar = np.ndim(ell) == 0
        self.ells = [int(ell)] if self.scalar else [int(e) for e in ell]
        self.q = 1.5 + q
        self.size = self.k.size
        self.delta = np.log(self.k[-1] / self.k[0]) / (self.size - 1)
        nfolds = (self.size * minfolds - 1).bit_length()
        self.padded_size = 2**nfolds
        npad = self.padded_size - self.size
        self.pad_in_left, self.pad_in_right = npad // 2, npad - npad // 2
        self.pad_out_left, self.pad_out_right = npad - npad // 2, npad // 2
        n = np.arange(-self.pad_in_left, self.size + self.pad_in_right)
        self.padded_k = self.k[0] * np.exp(n * self.delta)
        m = np.arange(0, self.padded_size // 2 + 1)
        self.s, self.u, self.prefactor, self.postfactor = [], [], [], []
        for ell in self.ells:
            if lowring:
                lnxy = self.delta / np.pi * np.angle(spherical_bessel_kernel(ell, self.q + 1j * np.pi / self.delta))
            else:
                lnxy = self.delta
            s = np.exp(lnxy - self.delta) / self.k[::-1]
            n_out = np.arange(-self.pad_out_left, self.size + self.pad_out_right)
            padded_s = s[0] * np.exp(n_out * self.delta)
            u = spherical_bessel_kernel(ell, self.q + 2j * np.pi / self.padded_size / self.delta * m)
            u = u * np.exp(-2j * np.pi * lnxy / self.padded_size / self.delta * m)
            u[-1] = u[-1].real  # Nyquist term real
            self.s.append(s)
            self.u.append(u)
            self.prefactor.append(self.padded_k**(3. - self.q) / (2. * np.pi)**1.5)
            self.postfactor.append(padded_s**(-self.q) * (-1.)**(ell // 2))
        self.s, self.u = np.array(self.s), np.array(self.u)
        self.prefactor, self.postfactor = np.array(self.prefactor), np.array(self.postfactor)

    def __call__(self, fun):
        fun = np.asarray(fun, dtype='f8')
        if self.scalar and fun.ndim == 1: fun = fun[None, :]
        lead = fun.shape[:-2]
        padded = np.zeros(lead + (len(self.ells), self.padded_size), dtype='f8')
        padded[..., self.pad_in_left:self.pad_in_left + self.size] = fun
        spec = np.fft.rfft(padded * self.prefactor, axis=-1) * self.u
        out = np.fft.hfft(spec, n=self.padded_size, axis=-1) / self.padded_size * self.postfactor
        out = out[..., self.pad_out_left:self.pad_out_left + self.size]
        if self.scalar:
            return self.s[0], out[..., 0, :]
        return self.s, out


def brute_force_xi(k, pk, ell, s):
    """Defining integral, restated as in reference desilike/theories/galaxy_clustering/base.py:163-168."""
    from scipy.special import spherical_jn
    k = np.asarray(k)
    lnk = np.log(k)
    w = np.empty_like(k)
    w[1:-1] = (lnk[2:] - lnk[:-2]) / 2.
    w[0] = (lnk[1] - lnk[0]) / 2.
    w[-1] = (lnk[-1] - lnk[-2]) / 2.
    return (-1.)**(ell // 2) / (2. * np.pi**2) * np.sum(k**3 * pk * w * spherical_jn(ell, np.asarray(s)[:, None] * k), axis=-1)
