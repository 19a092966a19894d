"""desilike_b200: B200-native (sm_100a CUDA) batched full-shape likelihood path of cosmodesi/desilike.

Scope: ONE hot path -- N parameter points through emulated / Kaiser / BAO theory -> window matrix -> Gaussian chi2 with
analytic marginalisation -> prior -- behind desilike's calling conventions.  See DESIGN.md / INTEGRATION.md.
"""
from ._capi import DesilikeB200Error, load as load_library
from .plan import Plan, launch_count, measure_fp64_peak, pinned_empty, sample_solved
from .lowering import Lowered, LoweringError

from .parameter import Parameter, ParameterCollection, Samples
from .theories import (ShapeFitPowerSpectrumTemplate, BAOPowerSpectrumTemplate, KaiserTracerPowerSpectrumMultipoles, KaiserTracerCorrelationFunctionMultipoles, EFTLikeKaiserTracerCorrelationFunctionMultipoles,
                       EFTLikeKaiserTracerPowerSpectrumMultipoles, LPTVelocileptorsTracerPowerSpectrumMultipoles, TableBiasTracerPowerSpectrumMultipoles, DampedBAOWigglesTracerPowerSpectrumMultipoles,
                       DampedBAOWigglesTracerCorrelationFunctionMultipoles)
from .emulators import Emulator, EmulatedCalculator
from .observables import TracerPowerSpectrumMultipolesObservable, TracerCorrelationFunctionMultipolesObservable
from .likelihoods import ObservablesGaussianLikelihood, SumLikelihood, PipelineError
from .base import vmap, status_to_errors
from .samplers import EnsembleSampler
from .adapters import spec_from_likelihood

__all__ = ['Parameter', 'ParameterCollection', 'Samples', 'ShapeFitPowerSpectrumTemplate', 'BAOPowerSpectrumTemplate',
           'KaiserTracerPowerSpectrumMultipoles', 'KaiserTracerCorrelationFunctionMultipoles', 'EFTLikeKaiserTracerCorrelationFunctionMultipoles', 'EFTLikeKaiserTracerPowerSpectrumMultipoles', 'LPTVelocileptorsTracerPowerSpectrumMultipoles', 'TableBiasTracerPowerSpectrumMultipoles', 'DampedBAOWigglesTracerPowerSpectrumMultipoles',
           'DampedBAOWigglesTracerCorrelationFunctionMultipoles', 'Emulator', 'EmulatedCalculator', 'TracerPowerSpectrumMultipolesObservable',
           'TracerCorrelationFunctionMultipolesObservable', 'ObservablesGaussianLikelihood', 'SumLikelihood', 'PipelineError', 'vmap', 'status_to_errors', 'EnsembleSampler', 'spec_from_likelihood',
           'Plan', 'sample_solved', 'Lowered', 'LoweringError', 'DesilikeB200Error', 'load_library', 'launch_count', 'measure_fp64_peak', 'pinned_empty']
__version__ = '0.1.0'
