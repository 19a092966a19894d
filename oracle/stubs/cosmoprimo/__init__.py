"""Empty-shell stand-in for the third-party `cosmoprimo` package (absent here, no network).

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own NumPy code path
(/root/reference/desilike, read-only) can be imported in the build container to
pin the oracle.  Nothing in the product imports it.
Only names are provided; the arithmetic that the reference delegates to cosmoprimo
(FFTLog, emulator predict) is supplied by the oracle restatements, which
are injected here so the reference's calculators can call them.
"""
from . import constants, fiducial
from .cosmology import Cosmology, CosmologyError
from .interpolator import PowerSpectrumInterpolator1D, PowerSpectrumInterpolator2D


class PowerSpectrumBAOFilter(object):
    """Returns the synthetic no-wiggle spectrum (oracle/synthetic.py) as the 'smoothed' P(k)."""

    def __init__(self, pk_interpolator=None, engine=None, cosmo=None, cosmo_fid=None, **kwargs):
        pass

    def __call__(self, pk_interpolator=None, cosmo=None, **kwargs):
        return self

    def smooth_pk_interpolator(self):
        from oracle import synthetic
        from .cosmology import _Interp1D
        return _Interp1D(synthetic.pknow_fid)


class PowerToCorrelation(object):
    """Routes to the oracle's FFTLog restatement (oracle/fftlog.py)."""

    def __init__(self, k, ell=0, q=0, lowring=True, **kwargs):
        from oracle.fftlog import PowerToCorrelation as _P2C
        self._impl = _P2C(k, ell=ell, q=q, lowring=lowring, **kwargs)

    def __call__(self, fun):
        return self._impl(fun)
