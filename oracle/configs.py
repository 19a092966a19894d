"""Seeded synthetic inputs for BASELINE.json configs C1-C5, as likelihood specs.

TEST / BENCH INFRASTRUCTURE.  The generators live product-side (desilike_b200/workloads.py: no reference arithmetic, ``bench.py`` builds its inputs
from them without importing this package); here they are re-exported with the ORACLE installed for the two steps that need a theory evaluation
(synthetic data vector = theory at a fiducial point; finite-difference Taylor coefficients of C2(i)), so that the CPU test suite and the reference
arm need no GPU.
"""
from desilike_b200 import workloads as _workloads
from desilike_b200.workloads import *   # noqa: F401,F403
from desilike_b200.workloads import _FID, _POINT_RANGES, _finish, _densify, _kaiser_taylor_engine, _random_taylor_engine   # noqa: F401

from . import pipeline as _pipeline

_workloads.set_hooks(theory_flat=_pipeline.observable_flat, theory_tables=_pipeline.theory_tables)
