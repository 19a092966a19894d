"""Plan extraction from an initialised *reference* likelihood: the ``spec_from_likelihood`` step of INTEGRATION.md.

A ``backend='cuda'`` branch of the reference's ``vmap`` (desilike/base.py:232-383) receives desilike's own objects.  This module reads
the constants of such a pipeline -- parameters and priors (reference parameter.py), window matrix / offset / shot noise of every
observable (observables/galaxy_clustering/window.py:207-457), template tables of the Kaiser theory (power_template.py:747-761,
full_shape.py:470-500), data vector and precision (likelihoods/base.py:530-656) -- into the likelihood *spec* that ``desilike_b200.Plan``
lowers (lowering.py).  Nothing of the reference is imported here: the objects are only read through the attributes cited below, so the
module works on whatever desilike installation the caller has.

Supported (everything else fails loudly): ``ObservablesGaussianLikelihood`` over ``TracerPowerSpectrumMultipolesObservable``s whose theory is
``KaiserTracerPowerSpectrumMultipoles`` / ``EFTLikeKaiserTracerPowerSpectrumMultipoles`` with a ``ShapeFitPowerSpectrumTemplate`` (direct
theory), including systematic templates, the 'cubic' transform and '.marg' / '.best' solved parameters ('.prec' parameters are folded into
the reference object's precision when it initialises, so a likelihood carrying them is rejected).  Tested against live reference objects in
tests/test_oracle_vs_reference.py (build container; the reference does not travel to the GPU box).
"""
import numpy as np

from .likelihoods import resolve_terms
from .theories import KaiserTracerPowerSpectrumMultipoles, EFTLikeKaiserTracerPowerSpectrumMultipoles

_THEORIES = {'KaiserTracerPowerSpectrumMultipoles': KaiserTracerPowerSpectrumMultipoles,
             'EFTLikeKaiserTracerPowerSpectrumMultipoles': EFTLikeKaiserTracerPowerSpectrumMultipoles}


def _prior(prior):
    """reference ParameterPrior (parameter.py:1900-2017) -> dict(dist, limits[, loc, scale])."""
    if prior is None: return None
    dist = getattr(prior, 'dist', 'uniform')
    out = dict(dist=dist, limits=tuple(float(x) for x in getattr(prior, 'limits', (-np.inf, np.inf))))
    if dist == 'norm': out.update(loc=float(prior.loc), scale=float(prior.scale))
    return out


def spec_from_likelihood(likelihood, solved_default='.marg'):
    """Initialised reference ``ObservablesGaussianLikelihood`` -> likelihood spec (``desilike_b200.Plan(spec)``)."""
    if not hasattr(likelihood, 'observables') or not hasattr(likelihood, 'precision'):
        raise NotImplementedError('spec_from_likelihood expects an ObservablesGaussianLikelihood (got {})'.format(type(likelihood).__name__))
    params = [p for p in likelihood.all_params if p.derived is not True]   # derived=True: outputs (loglikelihood, f_sqrt_Ap, ...), not inputs
    varied = [p.name for p in params if p.varied and not p.solved]
    solved = [p for p in params if p.solved]
    if any(str(p.derived).startswith('.prec') for p in solved):
        raise NotImplementedError("'.prec' parameters are already folded into the reference likelihood's precision at initialisation: build the spec before "
                                  "that step (desilike_b200.ObservablesGaussianLikelihood handles '.prec' itself)")
    fixed = {p.name: float(p.value) for p in params if not p.varied}
    defaults = {p.name: float(p.value) for p in params}
    spec = dict(params=varied, priors={p.name: _prior(p.prior) for p in params if p.name in varied},
                solved=[dict(name=p.name, value=float(p.value), prior=_prior(p.prior), marg=str(p.derived).replace('.auto', solved_default).startswith('.marg'),
                             mode='prec' if str(p.derived).startswith('.prec') else None) for p in solved],
                fiducial={name: defaults[name] for name in varied}, observables=[])
    for obs in likelihood.observables:
        w = getattr(obs, 'wmatrix', None)
        if w is None or not hasattr(w, 'kin'):
            raise NotImplementedError('observable {} is not a windowed power-spectrum observable'.format(type(obs).__name__))
        th = w.theory
        cls = _THEORIES.get(type(th).__name__)
        if cls is None: raise NotImplementedError('theory {} is not on the CUDA path through this adapter'.format(type(th).__name__))
        pt, tpl = th.pt, th.pt.template
        if type(tpl).__name__ != 'ShapeFitPowerSpectrumTemplate':
            raise NotImplementedError('template {} is not on the CUDA path through this adapter'.format(type(tpl).__name__))
        kin, ellsin = np.asarray(w.kin, dtype='f8'), tuple(w.ellsin)
        b1 = [p for p in th.all_params if p.basename == 'b1'][0]
        # bias-combine terms from the mirror class (full_shape.py:545-550, :628-634); only k, ells, nd and the parameter names are needed
        mirror = cls(template=None, k=kin, ells=ellsin, shotnoise=1. / float(th.nd), tracer=b1.namespace or None)
        if hasattr(th, 'kp'): mirror.kp = float(th.kp)
        terms = list(mirror.terms())
        st = getattr(w, 'systematic_templates', None)
        if st is not None:   # window.py:470-473: flatpower += sum_i syst_i template_i
            names = {p.basename: p.name for p in st.all_params}
            for name, template in st.templates.items():
                terms.append(dict(scale=1., factors=[(names.get(name, name), 1)], table=None, outvec=np.array(template, dtype='f8')))
        theory = dict(kind='kaiser', k=kin, ells=ellsin, nmu=len(pt.mu),
                      template=dict(kind='shapefit', k=np.array(tpl.k, dtype='f8'), pk_dd_fid=np.array(tpl.pk_dd_fid, dtype='f8'),
                                    pknow_dd_fid=np.array(tpl.pknow_dd_fid, dtype='f8'), f_fid=float(tpl.f_fid), a=float(tpl.a), kp=float(tpl.kp),
                                    apmode=str(tpl.apeffect.mode), eta=float(getattr(tpl, 'eta', 1. / 3.)), only_now=bool(getattr(tpl, 'only_now', False))),
                      params=[n for n in varied if n.split('.')[-1] in ('qpar', 'qper', 'qiso', 'qap', 'dm', 'dn', 'df', 'sigmapar', 'sigmaper')],
                      defaults={n.split('.')[-1]: v for n, v in defaults.items()})
        # window constants (window.py:459-473): flat = (W . ravel(P + N_in) + offset)[kmask] - N_out
        nin = len(ellsin) * kin.size
        W = None if w.matrix_full is None else np.array(w.matrix_full, dtype='f8')
        noise_in = np.repeat(np.asarray(w.shotnoisein, dtype='f8'), kin.size)
        offset = W.dot(noise_in) if W is not None else noise_in.copy()
        if getattr(w, 'offset', None) is not None: offset = offset + np.asarray(w.offset, dtype='f8')
        if getattr(w, 'kmask', None) is not None:
            W = (np.eye(nin) if W is None else W)[w.kmask]
            offset = offset[w.kmask]
        offset = offset - np.asarray(w.shotnoiseout, dtype='f8')
        spec['observables'].append(dict(name=getattr(obs, 'name', 'observable'), nin=nin, wmatrix=W, offset=offset, transform=getattr(obs, 'transform', None), theory=theory,
                                        terms=resolve_terms(terms, [p.name for p in solved], fixed), data=np.array(obs.flatdata, dtype='f8')))
    spec['precision'] = np.array(likelihood.precision, dtype='f8')
    if len(spec['observables']) == 1: spec['precision'] = [spec['precision']]
    return spec
