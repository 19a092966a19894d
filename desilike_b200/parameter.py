"""Minimal parameter containers mirroring the part of desilike/parameter.py the batched path needs:
name / value / fixed / prior / derived and name-based selection.  ``derived``: '.marg', '.best', '.auto', '.prec' mark analytically solved
parameters (reference parameter.py:659,864-866); any other string is an EXPRESSION of other parameters, '{b}**2' (reference parameter.py:795-808),
compiled for the device by desilike_b200/expressions.py.  Not a re-implementation of the reference's 3000-line module: no Samples algebra."""
import fnmatch

import numpy as np


class Parameter(object):

    def __init__(self, basename, value=None, fixed=False, prior=None, ref=None, derived=False, namespace='', latex=None, delta=None):
        self.basename, self.namespace = str(basename), str(namespace or '')
        self.prior = dict(prior) if prior is not None else None
        if self.prior is not None and 'dist' not in self.prior:
            self.prior['dist'] = 'uniform'
        self.ref = dict(ref) if ref is not None else None
        if value is None:   # reference default: centre of the ref (else prior) distribution
            value = 0.
            for dist in (self.ref, self.prior):
                if dist is not None:
                    if 'limits' in dist and dist.get('dist', 'uniform') == 'uniform' and np.all(np.isfinite(dist['limits'])): value = float(np.mean(dist['limits']))
                    elif 'loc' in dist: value = float(dist['loc'])
                    break
        self.value, self.fixed, self.derived, self.latex, self.delta = float(value), bool(fixed), derived, latex, delta

    @property
    def name(self):
        return '{}.{}'.format(self.namespace, self.basename) if self.namespace else self.basename

    @property
    def solved(self):
        """reference parameter.py:864-866."""
        return isinstance(self.derived, str) and self.derived.startswith('.') and not self.fixed

    @property
    def varied(self):
        return not self.fixed and (self.derived is False or self.solved)

    @property
    def depends(self):
        """Names of the parameters a derived-expression parameter is computed from (reference parameter.py ``Parameter.depends``); else []."""
        if isinstance(self.derived, str) and not self.derived.startswith('.'):
            from .expressions import depends
            return depends(self.derived)
        return []

    def update(self, **kwargs):
        for key, val in kwargs.items():
            if key == 'prior' and val is not None:
                val = dict(val)
                val.setdefault('dist', 'uniform')
            if not hasattr(self, key) and key not in ('name',):
                raise AttributeError('unknown parameter attribute {}'.format(key))
            setattr(self, key, val if key != 'value' else float(val))
        return self

    def clone(self, **kwargs):
        new = Parameter(self.basename, value=self.value, fixed=self.fixed, prior=self.prior, ref=self.ref, derived=self.derived,
                        namespace=self.namespace, latex=self.latex, delta=self.delta)
        return new.update(**kwargs)

    def __repr__(self):
        return 'Parameter({}, value={}, {})'.format(self.name, self.value, 'solved' if self.solved else 'varied' if self.varied else 'fixed')


class ParameterCollection(object):

    def __init__(self, params=(), on_new=None):
        self.data = []
        self._on_new = on_new   # owner hook: ``collection[name] = {...}`` with a new name registers the parameter with the owner (a likelihood)
        for p in params: self.set(p)

    def __setitem__(self, name, item):
        """``params['b'] = {'prior': {'limits': [0., 2.]}}`` (reference parameter.py:1855-1871): update an existing parameter or add a new one."""
        name = getattr(name, 'name', name)
        if isinstance(item, Parameter):
            if item.name != name: raise KeyError('Parameter {} must be indexed by name (incorrect {})'.format(item.name, name))
            new = name not in self
            self.set(item)
            if new and self._on_new is not None: self._on_new(item)
            return
        if name in self:
            self[name].update(**dict(item))
            return
        namespace, _, basename = str(name).rpartition('.')
        param = Parameter(basename, namespace=namespace, **dict(item))
        self.set(param)
        if self._on_new is not None: self._on_new(param)

    def set(self, param):
        for i, p in enumerate(self.data):
            if p.name == param.name:
                self.data[i] = param
                return
        self.data.append(param)

    def __iter__(self):
        return iter(self.data)

    def __len__(self):
        return len(self.data)

    def __contains__(self, name):
        name = getattr(name, 'name', name)
        return any(p.name == name for p in self.data)

    def __getitem__(self, name):
        if isinstance(name, int): return self.data[name]
        for p in self.data:
            if p.name == name: return p
        for p in self.data:   # basename lookup when unambiguous
            if p.basename == name: return p
        raise KeyError(name)

    def names(self, **kwargs):
        return [p.name for p in self.select(**kwargs)]

    def select(self, varied=None, fixed=None, solved=None, basename=None, name=None):
        out = []
        patterns = [basename] if isinstance(basename, str) else basename
        npatterns = [name] if isinstance(name, str) else name
        for p in self.data:
            if varied is not None and p.varied != varied: continue
            if fixed is not None and p.fixed != fixed: continue
            if solved is not None and p.solved != solved: continue
            if patterns is not None and not any(fnmatch.fnmatch(p.basename, pat) for pat in patterns): continue
            if npatterns is not None and not any(fnmatch.fnmatch(p.name, pat) for pat in npatterns): continue
            out.append(p)
        return ParameterCollection(out)

    def __add__(self, other):
        new = ParameterCollection(self.data)
        for p in other: new.set(p)
        return new

    def __repr__(self):
        return 'ParameterCollection({})'.format(self.names())


class ParameterArray(np.ndarray):
    """Array of a derived quantity over the batch.  With ``derivs`` (a list of tags: ``()`` = the value, ``(name1, name2)`` = second
    derivative w.r.t. two solved parameters) the LAST axis runs over the tags, as in the reference's ``ParameterArray``
    (desilike/parameter.py:407-448, :570-597): ``array[()]`` is the value, ``array[('a', 'b')]`` the Hessian entry (order-insensitive)."""

    def __new__(cls, value, derivs=None):
        obj = np.asarray(value, dtype='f8').view(cls)
        obj.derivs = None if derivs is None else [tuple(d) for d in derivs]
        return obj

    def __array_finalize__(self, obj):
        self.derivs = getattr(obj, 'derivs', None)

    def __getitem__(self, index):
        derivs = getattr(self, 'derivs', None)
        if derivs is not None and isinstance(index, tuple) and all(isinstance(i, str) for i in index):
            key = tuple(sorted(index))
            for i, d in enumerate(derivs):
                if tuple(sorted(d)) == key: return np.asarray(self)[..., i]
            raise KeyError('{} is not in computed derivatives: {}'.format(index, derivs))
        out = np.ndarray.__getitem__(self, index)
        if isinstance(out, ParameterArray): out.derivs = None   # a generic slice no longer spans the tag axis
        return out


class Samples(dict):
    """Derived outputs of a batched call: name -> array shaped like the input batch (stand-in for desilike's Samples)."""

    def __getitem__(self, name):
        return dict.__getitem__(self, getattr(name, 'name', name))
