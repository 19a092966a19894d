"""Synthetic fiducial inputs for the hot path (TEST / BENCH INFRASTRUCTURE; not product code).

The reference obtains ``pk_dd_fid``, ``pknow_dd_fid``, ``f_fid`` and ``rs_drag`` from
cosmoprimo/CLASS at initialisation (reference desilike/theories/galaxy_clustering/power_template.py:49-66,77-101).
Those are *constants* of the batched path, not part of its arithmetic, and cosmoprimo is
absent here, so SURVEY.md section 8(d) replaces them by these closed-form stand-ins.
"""
import numpy as np

F_FID = 0.8
RS_DRAG = 100.  # Mpc/h, sets the BAO wiggle period ~ 2 pi / rs
SIGMA8 = 0.8


def pknow_fid(k):
    """Smooth BBKS-like broadband, amplitude ~2e4 (Mpc/h)^3 at k = 0.02 h/Mpc."""
    k = np.asarray(k, dtype='f8')
    q = k / 0.2 * 1.45  # Gamma-like shape parameter
    T = np.log(1. + 2.34 * q) / (2.34 * q) * (1. + 3.89 * q + (16.1 * q)**2 + (5.46 * q)**3 + (6.71 * q)**4)**(-0.25)
    pk = k**0.965 * T**2
    q0 = 0.02 / 0.2 * 1.45
    T0 = np.log(1. + 2.34 * q0) / (2.34 * q0) * (1. + 3.89 * q0 + (16.1 * q0)**2 + (5.46 * q0)**3 + (6.71 * q0)**4)**(-0.25)
    return 2e4 * pk / (0.02**0.965 * T0**2)


def pk_fid(k):
    """Broadband times a damped BAO-like wiggle."""
    k = np.asarray(k, dtype='f8')
    return pknow_fid(k) * (1. + 0.05 * np.sin(k * 105.) * np.exp(-(k * 8.)**2))
