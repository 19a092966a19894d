"""Host-side mirror of the reference's Gaussian likelihood on the batched path.

  ObservablesGaussianLikelihood   reference likelihoods/base.py:506-664 (covariance -> precision :605-631, Hartlap :623-629,
                                  Percival :633-656, chi2 :658-660), result = loglikelihood + logprior (:242-245, :413)
  SumLikelihood                   reference likelihoods/base.py:715-732
Calling ``likelihood(**params)`` evaluates ONE point through the same CUDA path as a batch of one; ``vmap(likelihood)`` is the
batched entry (desilike_b200/base.py).
"""
import numpy as np

from .parameter import ParameterCollection, Samples
from .plan import Plan


class PipelineError(Exception):
    """reference desilike/base.py (unknown parameter name etc.)."""


def resolve_terms(terms, solved_names, fixed):
    """Theory terms with parameter factors -> spec terms: a factor that is a solved parameter (must enter linearly) becomes the term's
    ``solved`` tag, a fixed parameter is folded into the scale (terms with zero scale drop out), sampled parameters stay factors."""
    out = []
    for term in terms:
        term = dict(term)
        scale, factors, sname = float(term.get('scale', 1.)), [], term.get('solved')
        for name, power in term.get('factors', []):
            if name in solved_names:
                if power != 1 or sname is not None: raise ValueError('solved parameter {} must enter linearly'.format(name))
                sname = name
            elif name in fixed:
                scale *= fixed[name]**power
            else:
                factors.append((name, power))
        if scale == 0.: continue
        out.append(dict(term, scale=scale, factors=factors, solved=sname))
    return out


class BaseLikelihood(object):
    solved_default = '.marg'   # reference likelihoods/base.py:207

    def __init__(self):
        self._plan, self._spec, self._signature = None, None, None
        self._extra_params = ParameterCollection()   # parameters added through ``likelihood.all_params[name] = {...}`` (inputs of derived expressions)
        self.loglikelihood = self.logprior = None
        self.flattheory = self.flatdiff = None

    # ---- parameters ----
    @property
    def all_params(self):
        params = ParameterCollection()
        for obs in self.observables: params = params + obs.theory.all_params + obs.params
        if not hasattr(self, '_extra_params'): self._extra_params = ParameterCollection()
        params = params + self._extra_params
        params._on_new = self._extra_params.set
        return params

    def _params_signature(self):
        """Cheap fingerprint of everything a plan depends on in the parameters: a later ``Parameter.update(fixed=..., prior=..., value=..., derived=...)``
        rebuilds the plan at the next evaluation (the reference re-initialises its pipeline on parameter updates)."""
        return tuple((p.name, p.value, p.fixed, repr(p.prior), repr(p.derived)) for p in self.all_params)

    @property
    def varied_params(self):
        return ParameterCollection([p for p in self.all_params if p.varied and not p.solved])

    @property
    def solved_params(self):
        return ParameterCollection([p for p in self.all_params if p.solved])

    # ---- spec / plan ----
    def _invalidate(self):
        if self._plan is not None: self._plan.close()
        self._plan, self._spec = None, None

    def build_spec(self):
        all_params = self.all_params
        varied = [p.name for p in all_params if p.varied and not p.solved]
        solved = [p for p in all_params if p.solved]
        solved_names = [p.name for p in solved]
        # derived-expression parameters ('{b}**2', reference parameter.py:795-808): evaluated per point on the device; an expression may only use
        # sampled parameters and derived ones defined before it (fixed parameters are substituted by their values)
        exprs = [p for p in all_params if p.depends]
        fixed = {p.name: p.value for p in all_params if not p.varied and not p.depends}
        derived, known = [], set(varied)
        pending = list(exprs)
        while pending:
            progress = False
            for p in list(pending):
                expr, ok = p.derived, True
                for name in p.depends:
                    if name in fixed: expr = expr.replace('{' + name + '}', '({!r})'.format(float(fixed[name])))
                    elif name not in known: ok = False
                if ok:
                    derived.append(dict(name=p.name, expr=expr)); known.add(p.name); pending.remove(p); progress = True
            if not progress:
                raise PipelineError('derived parameters {} depend on unknown (or solved) parameters'.format([p.name for p in pending]))
        derived_names = [dp['name'] for dp in derived]
        spec = dict(params=varied, derived=derived, priors={p.name: p.prior for p in all_params if p.name in varied}, observables=[], precision=None,
                    solved=[dict(name=p.name, value=p.value, prior=p.prior, marg=p.derived.replace('.auto', self.solved_default).startswith('.marg'),
                                 mode='prec' if p.derived.startswith('.prec') else None) for p in solved],
                    fiducial={p.name: p.value for p in all_params if p.name in varied})   # '.prec': Jacobian at the default values (reference likelihoods/base.py:288-296)
        for obs in self.observables:
            ospec = obs.observable_spec()
            theory = obs.theory.theory_spec()
            theory['params'] = [n for n in theory['params'] if n in varied or n in derived_names]
            theory['defaults'] = dict(theory.get('defaults', {}))
            terms = resolve_terms(list(obs.theory.terms()) + list(obs.systematic_terms()), solved_names, fixed)
            ospec.update(theory=theory, terms=terms)
            spec['observables'].append(ospec)
        return spec

    def _finalize_spec(self, spec):
        return spec

    @property
    def plan(self):
        signature = self._params_signature()
        if self._plan is not None and getattr(self, '_signature', None) is None: self._signature = signature   # plan set from outside
        if self._plan is not None and signature != self._signature: self._invalidate()
        if self._plan is None:
            self._spec = self._finalize_spec(self.build_spec())
            self._plan = Plan(self._spec)
            self._signature = signature
        return self._plan

    @property
    def spec(self):
        self.plan
        return self._spec

    # ---- evaluation ----
    def _points(self, params, shape=None):
        plan = self.plan
        known = set(self.all_params.names())
        for name in params:
            if name not in known: raise PipelineError('input parameter {} is not one of parameters: {}'.format(name, sorted(known)))
        defaults = {p.name: p.value for p in self.all_params}
        cols = []
        n = None
        for name in plan.params:
            col = np.asarray(params[name], dtype='f8').ravel() if name in params else None
            if col is not None: n = col.size
            cols.append(col)
        n = n or 1
        return np.column_stack([col if col is not None else np.full(n, defaults[name]) for col, name in zip(cols, plan.params)])

    def __call__(self, *args, return_derived=False, **kwargs):
        params = args[0] if (len(args) == 1 and not kwargs) else kwargs
        if args and kwargs: raise PipelineError('could not interpret input args = {}, kwargs = {}'.format(args, kwargs))
        X = self._points({k: [v] for k, v in params.items()})
        res = self.plan.eval(X)
        self.loglikelihood, self.logprior = float(res['loglikelihood'][0]), float(res['logprior'][0])
        self.flattheory = self.plan.theory(X)[0]
        self.flatdiff = self.flattheory - self.flatdata
        if return_derived:
            derived = Samples(loglikelihood=res['loglikelihood'][0], logprior=res['logprior'][0])
            for i, name in enumerate(self.plan.solved_names): derived[name] = res['x_solved'][0, i]
            return float(res['logposterior'][0]), derived
        return float(res['logposterior'][0])


class ObservablesGaussianLikelihood(BaseLikelihood):

    def __init__(self, observables, covariance=None, scale_covariance=1., correct_covariance='hartlap-percival2014', precision=None, name=None):
        super().__init__()
        if not isinstance(observables, (list, tuple)): observables = [observables]
        self.observables, self.name = list(observables), name
        self.nobs = None
        if isinstance(correct_covariance, dict):
            self.nobs = correct_covariance.get('nobs')
            correct_covariance = correct_covariance['correction']
        self.correct_covariance = correct_covariance
        sizes = [obs.size for obs in self.observables]
        n = sum(sizes)
        self._groups = None
        if covariance is None and precision is None:
            if not all(obs.covariance is not None for obs in self.observables):
                raise ValueError('Observables must have their own covariance if global covariance or precision matrix not provided')
            self._groups = [[i] for i in range(len(self.observables))]
            self.covariance = [np.atleast_2d(obs.covariance) * scale_covariance for obs in self.observables]
            self.precision = [np.linalg.inv(c) for c in self.covariance]
        elif precision is not None:
            self.covariance = None
            self.precision = np.array(precision, dtype='f8') / scale_covariance
            if self.precision.shape[0] != n:
                raise ValueError('based on provided observables, precision expected to be of size {:d}, but found {}'.format(n, self.precision.shape))
        else:
            self.covariance = np.atleast_2d(np.array(covariance, dtype='f8')) * scale_covariance
            if self.covariance.shape != (n, n):
                raise ValueError('based on provided observables, covariance expected to be a matrix of shape ({0:d}, {0:d}), but found {1}'.format(n, self.covariance.shape))
            self.precision = np.linalg.inv(self.covariance)
        if self.nobs is not None and 'hartlap' in self.correct_covariance:   # reference likelihoods/base.py:623-629
            self.hartlap2007_factor = (self.nobs - n - 2.) / (self.nobs - 1.)
            self.precision = [p * self.hartlap2007_factor for p in self.precision] if isinstance(self.precision, list) else self.precision * self.hartlap2007_factor

    def _finalize_spec(self, spec):
        precision = self.precision
        if self.nobs is not None and 'percival' in self.correct_covariance:   # reference likelihoods/base.py:633-656
            nbins = sum(obs.size for obs in self.observables)
            A = 2. / (self.nobs - nbins - 1.) / (self.nobs - nbins - 4.)
            B = (self.nobs - nbins - 2.) / (self.nobs - nbins - 1.) / (self.nobs - nbins - 4.)
            # varied (sampled + solved) parameters reachable from THIS likelihood's observables (reference likelihoods/base.py:641-652): in a
            # SumLikelihood the spec carries the union of all members' parameters
            used = set()
            for ospec in spec['observables']:
                used |= set(ospec['theory'].get('params', []))
                for term in ospec['terms']:
                    used |= {n for n, _ in term.get('factors', [])}
                    if term.get('solved') is not None: used.add(term['solved'])
            for dp in spec.get('derived', []):   # a derived-expression parameter stands for the sampled parameters it is computed from
                if dp['name'] in used:
                    from .expressions import depends
                    used |= set(depends(dp['expr']))
            varied_names = set(spec['params']) | {s['name'] for s in spec['solved']}
            nparams = len(used & varied_names)
            self.percival2014_factor = (1 + B * (nbins - nparams)) / (1 + A + B * (nparams + 1))
            precision = [p / self.percival2014_factor for p in precision] if isinstance(precision, list) else precision / self.percival2014_factor
        spec['precision'] = precision
        if isinstance(precision, list): spec['prec_groups'] = [list(range(len(precision)))]   # the blocks of ONE likelihood: '.prec' updates them jointly
        # synthetic data generated from the theory at given parameters (reference power_spectrum.py:95-97)
        pending = [obs for obs in self.observables if obs.flatdata is None]
        for obs, ospec in zip(self.observables, spec['observables']):
            ospec['data'] = obs.flatdata if obs.flatdata is not None else np.zeros(obs.size)
        if pending:
            tmp = Plan(spec)
            defaults = {p.name: p.value for p in self.all_params}
            start = 0
            for obs, ospec in zip(self.observables, spec['observables']):
                if obs.flatdata is None:
                    vals = dict(defaults); vals.update(obs._data_params or {})
                    X = np.array([[vals[name] for name in tmp.params]])
                    obs.flatdata = tmp.theory(X)[0, start:start + obs.size].copy()
                    ospec['data'] = obs.flatdata
                start += obs.size
            tmp.close()
        self.flatdata = np.concatenate([obs.flatdata for obs in self.observables])
        return spec


class SumLikelihood(BaseLikelihood):
    """Sum of independent Gaussian likelihoods (reference likelihoods/base.py:715-732): one plan, one block per member."""

    def __init__(self, likelihoods):
        super().__init__()
        self.likelihoods = list(likelihoods)
        self.observables = [obs for like in self.likelihoods for obs in like.observables]

    def _finalize_spec(self, spec):
        precisions, groups, offset = [], [], 0
        for like in self.likelihoods:
            nobs = len(like.observables)
            sub = dict(spec, observables=spec['observables'][offset:offset + nobs])
            sub.pop('prec_groups', None)
            sub = like._finalize_spec(sub)
            first = len(precisions)
            if not isinstance(sub['precision'], list):
                precisions.append(sub['precision'])   # one block for the member (joint over its observables)
            else:
                precisions += list(sub['precision'])
            groups.append(list(range(first, len(precisions))))   # '.prec' pre-marginalisation acts per member likelihood
            offset += nobs
        spec['precision'], spec['prec_groups'] = precisions, groups
        self.flatdata = np.concatenate([obs.flatdata for obs in self.observables])
        return spec
