"""Synthetic stand-in cosmology (see package docstring): closed-form P(k), distances and rs_drag."""
import numpy as np


class CosmologyError(Exception):
    pass


class _Interp1D(object):
    def __init__(self, fun):
        self._fun = fun

    def __call__(self, k):
        return self._fun(np.asarray(k, dtype='f8'))

    def sigma_r(self, r):
        k = np.logspace(-5, 2, 4096)
        x = k * r
        w = 3. * (np.sin(x) - x * np.cos(x)) / x**3
        return np.sqrt(np.trapezoid(self(k) * w**2 * k**3 / (2. * np.pi**2), np.log(k)))


class _Interp2D(object):
    def __init__(self, fun):
        self._fun = fun

    def to_1d(self, z=None, **kwargs):
        return _Interp1D(self._fun)


class _Fourier(object):
    def sigma8_z(self, z, of='delta_cb'):
        from oracle import synthetic
        return synthetic.SIGMA8 * (synthetic.F_FID if of.startswith('theta') else 1.)

    def pk_interpolator(self, of='delta_cb', **kwargs):
        from oracle import synthetic
        if isinstance(of, str) and of.startswith('theta'):
            return _Interp2D(lambda k: synthetic.F_FID**2 * synthetic.pk_fid(k))
        return _Interp2D(synthetic.pk_fid)


class Cosmology(object):
    n_s = 0.965
    h = 0.7

    def __init__(self, *args, **kwargs):
        self._kwargs = kwargs

    @property
    def rs_drag(self):
        from oracle import synthetic
        return synthetic.RS_DRAG

    def efunc(self, z):
        z = np.asarray(z, dtype='f8')
        return np.sqrt(0.3 * (1. + z)**3 + 0.7)

    def comoving_angular_distance(self, z):
        z = np.asarray(z, dtype='f8')
        zz = np.linspace(0., 1., 513)[:, None] * z
        return 2997.92458 * np.trapezoid(1. / self.efunc(zz), zz, axis=0)

    def get_fourier(self, *args, **kwargs):
        return _Fourier()

    def __getstate__(self):
        return {}


class BaseEngine(object):
    pass


class BaseSection(object):
    pass


def find_conflicts(*args, **kwargs):
    return []
