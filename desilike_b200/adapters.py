"""Plan extraction from an initialised *reference* likelihood: the ``spec_from_likelihood`` step of INTEGRATION.md.

A ``backend='cuda'`` branch of the reference's ``vmap`` (desilike/base.py:232-383) receives desilike's own objects.  This module reads
the constants of such a pipeline -- parameters and priors (reference parameter.py), window matrix / offset / shot noise of every
observable (observables/galaxy_clustering/window.py:207-457), template tables of the Kaiser theory (power_template.py:747-761,
full_shape.py:470-500), data vector and precision (likelihoods/base.py:530-656) -- into the likelihood *spec* that ``desilike_b200.Plan``
lowers (lowering.py).  Nothing of the reference is imported here: the objects are only read through the attributes cited below, so the
module works on whatever desilike installation the caller has.

Supported (everything else fails loudly): ``ObservablesGaussianLikelihood`` over ``TracerPowerSpectrumMultipolesObservable``s whose theory is
``KaiserTracerPowerSpectrumMultipoles`` / ``EFTLikeKaiserTracerPowerSpectrumMultipoles`` (and their ``...CorrelationFunctionMultipoles`` forms) with a
``ShapeFitPowerSpectrumTemplate`` (direct theory) or with an EMULATED ``pt`` (the reference's ``EmulatedCalculator``: Taylor / MLP engines), and over power-spectrum / correlation-function observables of the damped-BAO tracers (``DampedBAOWigglesTracer*Multipoles``, power-law
broadband), including systematic templates, the 'cubic' transform and '.marg' / '.best' solved parameters ('.prec' parameters are folded into
the reference object's precision when it initialises, so a likelihood carrying them is rejected).  Tested against live reference objects in
tests/test_oracle_vs_reference.py (build container; the reference does not travel to the GPU box).
"""
import numpy as np

from .likelihoods import resolve_terms
from .theories import (KaiserTracerPowerSpectrumMultipoles, EFTLikeKaiserTracerPowerSpectrumMultipoles, KaiserTracerCorrelationFunctionMultipoles,
                       EFTLikeKaiserTracerCorrelationFunctionMultipoles)

_THEORIES = {'KaiserTracerPowerSpectrumMultipoles': KaiserTracerPowerSpectrumMultipoles,
             'EFTLikeKaiserTracerPowerSpectrumMultipoles': EFTLikeKaiserTracerPowerSpectrumMultipoles,
             'KaiserTracerCorrelationFunctionMultipoles': KaiserTracerCorrelationFunctionMultipoles,
             'EFTLikeKaiserTracerCorrelationFunctionMultipoles': EFTLikeKaiserTracerCorrelationFunctionMultipoles}


def _own_params(calculator):
    """Parameters of a reference calculator and of everything it requires, under their pipeline names (namespaces applied).

    NOT ``calculator.all_params``: on a sub-calculator of a likelihood that property builds a NEW pipeline for the sub-calculator (reference
    base.py:949-960, :1302-1305), which re-initialises the calculators it shares with the likelihood; the likelihood then rebuilds its own pipeline
    and every parameter update made on ``likelihood.all_params`` (``derived='.marg'``, priors, fixed flags) is lost -- extracting a plan must not
    change the object it reads.  Each calculator's ``runtime_info.params`` (base.py:915-921) and ``runtime_info._requires`` (:962-972) are plain
    reads once the pipeline is initialised.  Objects without ``runtime_info`` (mirror classes, test doubles) list their own ``all_params``."""
    seen, out, stack = set(), [], [calculator]
    while stack:
        calc = stack.pop()
        if id(calc) in seen: continue
        seen.add(id(calc))
        info = getattr(calc, 'runtime_info', None)
        if info is None or not hasattr(info, 'params'):
            out += list(calc.all_params)
            continue
        out += list(info.params)
        requires = getattr(info, '_requires', None)
        stack += list(requires if requires is not None else info.requires)
    names, unique = set(), []
    for p in out:
        if p.name not in names: names.add(p.name); unique.append(p)
    return unique


def _prior(prior):
    """reference ParameterPrior (parameter.py:1900-2017) -> dict(dist, limits[, loc, scale])."""
    if prior is None: return None
    dist = getattr(prior, 'dist', 'uniform')
    out = dict(dist=dist, limits=tuple(float(x) for x in getattr(prior, 'limits', (-np.inf, np.inf))))
    if dist == 'norm': out.update(loc=float(prior.loc), scale=float(prior.scale))
    return out


def spec_from_likelihood(likelihood, solved_default='.marg', all_params=None):
    """Initialised reference ``ObservablesGaussianLikelihood`` (or ``SumLikelihood`` of them) -> likelihood spec (``desilike_b200.Plan(spec)``).
    ``all_params``: parameter collection to read values / priors / solved tags from (default: the likelihood's own)."""
    if hasattr(likelihood, 'likelihoods'):   # SumLikelihood (reference likelihoods/base.py:715-732): independent members, one precision block each;
        # the parameters of the SUM's pipeline are authoritative (members configured separately may carry different priors for a shared name)
        members = [spec_from_likelihood(like, solved_default=solved_default, all_params=likelihood.all_params) for like in likelihood.likelihoods]
        spec = dict(members[0], observables=[], precision=[])
        for m in members:
            spec['observables'] += m['observables']
            spec['precision'] += list(m['precision']) if isinstance(m['precision'], list) else [m['precision']]
        return spec
    if not hasattr(likelihood, 'observables') or not hasattr(likelihood, 'precision'):
        raise NotImplementedError('spec_from_likelihood expects an ObservablesGaussianLikelihood (got {})'.format(type(likelihood).__name__))
    params = [p for p in (likelihood.all_params if all_params is None else all_params) if p.derived is not True]   # derived=True: outputs, not inputs
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
    solved_names = [p.name for p in solved]
    for obs in likelihood.observables:
        w = getattr(obs, 'wmatrix', None)
        if w is None or not (hasattr(w, 'kin') or hasattr(w, 'sin')):
            raise NotImplementedError('observable {} is not a windowed two-point observable'.format(type(obs).__name__))
        th = w.theory
        tname = type(th).__name__
        is_xi = hasattr(w, 'sin')
        xin, ellsin = np.asarray(w.sin if is_xi else w.kin, dtype='f8'), tuple(w.ellsin)
        nin = len(ellsin) * xin.size
        # the kernel slots of a theory are resolved by BASENAME: restrict names and defaults to the theory's OWN parameters (a SumLikelihood with
        # namespaced per-tracer parameters LRG.b1 / ELG.b1 must not drive one tracer's kernel with another tracer's values)
        own_params = _own_params(th)
        own = [p.name for p in own_params]
        basenames = {}
        for n in own:
            if n not in defaults: continue
            b = n.split('.')[-1]
            if b in basenames and n in varied and any(m in varied and m != n and m.split('.')[-1] == b for m in own):
                raise NotImplementedError('theory {} has two varied parameters with basename {!r}: kernel slots are resolved by basename'.format(tname, b))
            basenames[b] = defaults[n]

        def own_varied(slots):
            return [n for n in varied if n in own and n.split('.')[-1] in slots]
        if tname in _THEORIES and is_xi == tname.endswith('CorrelationFunctionMultipoles') and getattr(th.pt, 'emulator', None) is not None:
            # EMULATED perturbation theory: ``theory.pt`` is the reference's EmulatedCalculator (emulators/__init__.py:394-417); its emulator's state
            # (``Emulator.__getstate__``, :210-216 -- the dict the emulator files hold: engines{center, powers, derivatives | model_operations},
            # fixed{k, ells}, defaults) is read into the mirror emulator, and the mirror tracer class supplies the theory spec and the bias-combine terms
            from .emulators import Emulator, EmulatedCalculator
            emu = th.pt.emulator
            mirror_pt = EmulatedCalculator(Emulator(dict(emu.__getstate__())))
            b1 = [p for p in own_params if p.basename == 'b1'][0]
            if is_xi:
                mirror = _THEORIES[tname](pt=mirror_pt, s=xin, ells=ellsin, interp_order=int(th.interp_order), tracer=b1.namespace or None)
            else:
                mirror = _THEORIES[tname](pt=mirror_pt, k=xin, ells=ellsin, shotnoise=1. / float(th.nd), tracer=b1.namespace or None)
            if hasattr(th, 'kp'): mirror.kp = float(th.kp)
            elif hasattr(getattr(th, 'power', None), 'kp'): mirror.kp = float(th.power.kp)
            terms = list(mirror.terms())
            theory = dict(mirror.theory_spec())
            # emulator inputs that the likelihood does not vary sit at the likelihood's CURRENT values, not at the emulator's stored defaults
            theory['defaults'] = {**dict(theory.get('defaults', {})), **{p_: basenames[p_.split('.')[-1]] for p_ in theory['params'] if p_.split('.')[-1] in basenames}}
        elif tname in _THEORIES and is_xi == tname.endswith('CorrelationFunctionMultipoles'):
            pt, tpl = th.pt, th.pt.template
            if type(tpl).__name__ != 'ShapeFitPowerSpectrumTemplate':
                raise NotImplementedError('template {} is not on the CUDA path through this adapter'.format(type(tpl).__name__))
            b1 = [p for p in own_params if p.basename == 'b1'][0]
            # bias-combine terms from the mirror class (full_shape.py:545-550, :628-634); only k, ells, nd and the parameter names are needed
            if is_xi:   # Hankel transforms of the Kaiser tracer P_ell on kin (full_shape.py:336-365, :553-574)
                mirror = _THEORIES[tname](template=None, s=xin, ells=ellsin, interp_order=int(th.interp_order), tracer=b1.namespace or None)
            else:
                mirror = _THEORIES[tname](template=None, k=xin, ells=ellsin, shotnoise=1. / float(th.nd), tracer=b1.namespace or None)
            if hasattr(th, 'kp'): mirror.kp = float(th.kp)
            elif hasattr(getattr(th, 'power', None), 'kp'): mirror.kp = float(th.power.kp)   # correlation function: the EFT-like pivot lives on its power spectrum
            terms = list(mirror.terms())
            theory = dict(kind='kaiser', k=np.array(th.kin, dtype='f8') if is_xi else xin, ells=ellsin, nmu=len(pt.mu),
                          template=dict(kind='shapefit', k=np.array(tpl.k, dtype='f8'), pk_dd_fid=np.array(tpl.pk_dd_fid, dtype='f8'),
                                        pknow_dd_fid=np.array(tpl.pknow_dd_fid, dtype='f8'), f_fid=float(tpl.f_fid), a=float(tpl.a), kp=float(tpl.kp),
                                        apmode=str(tpl.apeffect.mode), eta=float(getattr(tpl, 'eta', 1. / 3.)), only_now=bool(getattr(tpl, 'only_now', False))),
                          params=own_varied(('qpar', 'qper', 'qiso', 'qap', 'dm', 'dn', 'df', 'sigmapar', 'sigmaper')), defaults=basenames)
            if is_xi: theory.update(s=xin, kin=np.array(th.kin, dtype='f8'), interp_order=int(th.interp_order))
        elif tname in ('DampedBAOWigglesTracerPowerSpectrumMultipoles', 'DampedBAOWigglesTracerCorrelationFunctionMultipoles') and is_xi == tname.endswith('CorrelationFunctionMultipoles'):
            # damped BAO (bao.py:117-151) + power-law broadband in x / pivot (bao.py:467-476, :527-534; :849-853, :905)
            pt = th.pt
            emulated = getattr(pt, 'emulator', None) is not None
            tpl = None if emulated else pt.template
            if str(getattr(th, 'broadband', 'power')) != 'power': raise NotImplementedError('only the power-law broadband is on the CUDA path')
            pivot = float(th.sp if is_xi else th.kp)
            names = {p.basename: p.name for p in own_params}
            terms = [dict(scale=1., factors=[], table='corr' if is_xi else 'power')]
            for ill, ell in enumerate(ellsin):
                for name, power in th.broadband_orders.get(ell, {}).items():
                    colvec = np.zeros((len(ellsin), xin.size)); colvec[ill] = (xin / pivot)**power
                    terms.append(dict(scale=1., factors=[(names.get(name, name), 1)], table=None, colvec=colvec.ravel()))
            if emulated:   # theory.pt = the reference's EmulatedCalculator of the damped-BAO P_ell ('power' engine on k, or on kin behind get_corr): state -> mirror emulator
                from .emulators import Emulator, EmulatedCalculator
                mirror_pt = EmulatedCalculator(Emulator(dict(pt.emulator.__getstate__())))
                if is_xi:   # the 'power' engine lives on kin (reference bao.py:820-833): the mirror class folds the P -> xi operator into its output map
                    from .theories import DampedBAOWigglesTracerCorrelationFunctionMultipoles
                    theory = dict(DampedBAOWigglesTracerCorrelationFunctionMultipoles(pt=mirror_pt, s=xin, ells=ellsin).theory_spec())
                else:
                    if len(mirror_pt.k) != xin.size or not np.allclose(mirror_pt.k, xin) or tuple(mirror_pt.ells) != ellsin:
                        raise ValueError('the emulator was built on a different (k, ells) grid than the one requested by the window / observable')
                    theory = dict(mirror_pt.theory_spec())
                    if 'power' not in theory['tables']: raise ValueError("the emulator has no 'power' engine (engines: {})".format(sorted(theory['tables'])))
                theory['defaults'] = {**dict(theory.get('defaults', {})), **{p_: basenames[p_.split('.')[-1]] for p_ in theory['params'] if p_.split('.')[-1] in basenames}}
            else:
                theory = dict(kind='bao_xi' if is_xi else 'bao_pk', ells=ellsin, nmu=len(pt.mu), mode=str(pt.mode), smoothing_radius=float(pt.smoothing_radius),
                          template=dict(kind='bao', k=np.array(tpl.k, dtype='f8'), pk_dd_fid=np.array(tpl.pk_dd_fid, dtype='f8'), pknow_dd_fid=np.array(tpl.pknow_dd_fid, dtype='f8'),
                                        f_fid=float(tpl.f_fid), apmode=str(tpl.apeffect.mode), eta=float(getattr(tpl, 'eta', 1. / 3.))),
                          params=own_varied(('qpar', 'qper', 'qiso', 'qap', 'df', 'b1', 'dbeta', 'sigmas', 'sigmapar', 'sigmaper')), defaults=basenames)
            if emulated: pass
            elif is_xi: theory.update(s=xin, kin=np.array(th.kin, dtype='f8') if hasattr(th, 'kin') else None, interp_order=int(getattr(th, 'interp_order', 1)))
            else: theory.update(k=xin)
        else:
            raise NotImplementedError('theory {} is not on the CUDA path through this adapter'.format(tname))
        st = getattr(w, 'systematic_templates', None)
        if st is not None:   # window.py:470-473 / :732-733: flat += sum_i syst_i template_i
            names = {p.basename: p.name for p in _own_params(st)}
            for name, template in st.templates.items():
                terms.append(dict(scale=1., factors=[(names.get(name, name), 1)], table=None, outvec=np.array(template, dtype='f8')))
        # window constants -> one matrix + offset
        W = None
        if is_xi:   # window.py:717-728: flat = (matrix_full . ravel(sum_l' matrix_diag[:, l'] theory[l']) + offset)[smask]
            md = getattr(w, 'matrix_diag', None)
            if md is not None:
                md = np.asarray(md, dtype='f8')
                ns = xin.size
                W = np.zeros((md.shape[0] * ns, md.shape[1] * ns))
                for lo in range(md.shape[0]):
                    for li in range(md.shape[1]): W[lo * ns:(lo + 1) * ns, li * ns:(li + 1) * ns] = np.diag(md[lo, li])
            if w.matrix_full is not None: W = np.array(w.matrix_full, dtype='f8') if W is None else np.asarray(w.matrix_full, dtype='f8').dot(W)
            offset = None if getattr(w, 'offset', None) is None else np.array(w.offset, dtype='f8')
            mask = getattr(w, 'smask', None)
            if mask is not None:
                W = (np.eye(nin) if W is None else W)[mask]
                if offset is not None: offset = offset[mask]
        else:       # window.py:459-473: flat = (W . ravel(P + N_in) + offset)[kmask] - N_out
            W = None if w.matrix_full is None else np.array(w.matrix_full, dtype='f8')
            noise_in = np.repeat(np.asarray(w.shotnoisein, dtype='f8'), xin.size)
            offset = W.dot(noise_in) if W is not None else noise_in.copy()
            if getattr(w, 'offset', None) is not None: offset = offset + np.asarray(w.offset, dtype='f8')
            if getattr(w, 'kmask', None) is not None:
                W = (np.eye(nin) if W is None else W)[w.kmask]
                offset = offset[w.kmask]
            offset = offset - np.asarray(w.shotnoiseout, dtype='f8')
        spec['observables'].append(dict(name=getattr(obs, 'name', 'observable'), nin=nin, wmatrix=W, offset=offset, transform=getattr(obs, 'transform', None), theory=theory,
                                        terms=resolve_terms(terms, solved_names, fixed), data=np.array(obs.flatdata, dtype='f8')))
    spec['precision'] = np.array(likelihood.precision, dtype='f8')
    if len(spec['observables']) == 1: spec['precision'] = [spec['precision']]
    return spec
