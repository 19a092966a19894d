"""Plan handle: owns the device constants of one likelihood and evaluates batches through the C-ABI."""
import ctypes

import numpy as np

from . import _capi
from .lowering import Lowered


def _is_torch_cuda(x):
    return hasattr(x, 'is_cuda') and x.is_cuda


class Plan(object):
    """
    Device-resident, immutable evaluation plan of a likelihood spec.

    ``plan.eval(X)`` replaces ``vmap(likelihood)(params)`` of the reference (desilike/base.py:232-273 driving
    likelihoods/base.py:242-245,314-413) for a batch ``X[N, d]`` (columns ordered as ``plan.params``).
    NumPy input -> host buffers (staged by the library, NumPy outputs); CUDA torch tensor input -> used in place on the
    current torch stream, torch outputs, no host round trip.
    """

    def __init__(self, spec, device=None, with_theory=True, reduce_rank=True, merge_terms=True, mlp_slices=0, qform='auto'):
        self._lib = _capi.load()
        if self._lib.dlb_device_count() <= 0:
            raise _capi.DesilikeB200Error('no CUDA device visible: desilike_b200 has no CPU fallback')
        if device is not None:
            import torch
            torch.cuda.set_device(device)
        spec = premarginalize_precision(spec)   # '.prec' parameters (plan-time, reference likelihoods/base.py:257-312)
        self.spec = spec
        # reduce_rank: blocks without solved parameters whose data vector is longer than their feature count are evaluated in the
        # row space of their term matrices (lowering.py: QR at plan time) -- same chi2, fewer columns
        # merge_terms: the tracer combine b1^2 dd + 2 b1 dt + tt of a direct Kaiser theory is formed inside the Kaiser kernel (lowering.py)
        # mlp_slices (opt-in, 2 .. 8): hidden layers of MLP emulators with fan-in >= 64 as FP64-grade GEMMs on the tcgen05 int8 tensor cores
        # (csrc/ozaki.cu); 0 keeps the FP64 DMMA kernels
        # qform: blocks with solved parameters as quadratic forms of their features ('auto': when that multiplies fewer columns; True / False)
        self.lowered = Lowered(spec, with_plain=with_theory, reduce_rank=reduce_rank, merge_terms=merge_terms, mlp_slices=mlp_slices, qform=qform)
        self.params = list(self.lowered.input_names)   # columns of the batch (derived-expression parameters are appended on the device)
        self.d, self.ns, self.ntheory = self.lowered.d_in, self.lowered.ns, self.lowered.ntheory
        self.solved_names = [s['name'] for s in spec.get('solved', [])]
        handle = ctypes.c_void_p()
        _capi.check(self._lib.dlb_plan_create(ctypes.byref(self.lowered.desc), ctypes.byref(handle)))
        self._handle = handle
        flags = self._lib.dlb_plan_flags(handle)
        # which kernels the plan runs (include/desilike_b200.h: dlb_plan_flags)
        self.kernels = {name: bool(flags >> bit & 1) for bit, name in enumerate(['resident_taylor', 'two_stage', 'fused_mlp', 'int8_layers', 'int8_contraction', 'quadratic_form'])}
        # plan order -> spec order of the solved parameters
        self._unperm = np.argsort(self.lowered.solved_order) if self.ns else np.zeros(0, dtype=int)
        self._identity_perm = bool(np.array_equal(self._unperm, np.arange(self.ns)))   # all '.marg' (or all '.best'): plan order == spec order

    def _theory_companion(self):
        """Plans whose dense precision was split over several passes carry no un-whitened matrices: theory vectors do not depend on the precision,
        so they come from a companion plan of the same spec with a unit diagonal precision per observable (built on first use)."""
        if getattr(self, '_companion', None) is None:
            spec = dict(self.spec, precision=[np.ones(len(o['data'])) for o in self.spec['observables']])
            spec.pop('prec_groups', None)
            self._companion = Plan(spec, with_theory=True)
        return self._companion

    def close(self):
        if getattr(self, '_companion', None) is not None:
            self._companion.close()
            self._companion = None
        if getattr(self, '_handle', None):
            self._lib.dlb_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    # ---- evaluation --------------------------------------------------------------------------------------------------
    def eval(self, X, return_solved=True, return_status=True, out=None, return_hessian=False):
        """-> dict(logposterior, loglikelihood, logprior[, x_solved (spec order), status, hessian]).

        ``return_hessian`` (NumPy input): ``hessian`` [N, ns, ns] = Hessian of the log-posterior in the solved parameters (spec order),
        the reference's ``posterior_hessian`` (likelihoods/base.py:187); ``sample_solved`` draws the marginalised parameters from it."""
        if _is_torch_cuda(X):
            return self._eval_torch(X, return_solved, return_status, out)
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.d:
            raise ValueError('expected params of shape (N, {:d}), got {}'.format(self.d, X.shape))
        N = X.shape[0]
        if out is not None:   # caller-provided (e.g. pinned) host buffers
            res = {key: out[key] for key in ['logposterior', 'loglikelihood', 'logprior']}
            xs, st = out.get('x_solved_plan_order'), out.get('status')
        else:
            res = dict(logposterior=np.empty(N), loglikelihood=np.empty(N), logprior=np.empty(N))
            xs = np.empty((N, self.ns)) if (return_solved and self.ns) else None
            st = np.empty(N, dtype=np.int32) if return_status else None
        hs = np.empty((N, self.ns * (self.ns + 1) // 2)) if (return_hessian and self.ns) else None
        _capi.check(self._lib.dlb_eval_batch_hessian(self._handle, X.ctypes.data, N, res['logposterior'].ctypes.data, res['loglikelihood'].ctypes.data,
                                                     res['logprior'].ctypes.data, xs.ctypes.data if xs is not None else None,
                                                     hs.ctypes.data if hs is not None else None, st.ctypes.data if st is not None else None, None))
        if hs is not None:
            iu = np.triu_indices(self.ns)
            H = np.zeros((N, self.ns, self.ns))
            H[:, iu[0], iu[1]] = hs
            H[:, iu[1], iu[0]] = hs
            res['hessian'] = -H[:, self._unperm][:, :, self._unperm]
        if xs is not None: res['x_solved'] = xs if self._identity_perm else xs[:, self._unperm]
        if st is not None: res['status'] = st
        return res

    def _eval_torch(self, X, return_solved, return_status, out):
        import torch
        if X.dtype != torch.float64 or not X.is_contiguous() or X.dim() != 2 or X.shape[1] != self.d:
            raise ValueError('expected a contiguous float64 CUDA tensor of shape (N, {:d})'.format(self.d))
        N = X.shape[0]
        own = out is None
        if out is None:
            out = dict(logposterior=torch.empty(N, dtype=torch.float64, device=X.device), loglikelihood=torch.empty(N, dtype=torch.float64, device=X.device),
                       logprior=torch.empty(N, dtype=torch.float64, device=X.device))
            if return_solved and self.ns: out['x_solved_plan_order'] = torch.empty((N, self.ns), dtype=torch.float64, device=X.device)
            if return_status: out['status'] = torch.empty(N, dtype=torch.int32, device=X.device)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        ptr = lambda key: out[key].data_ptr() if key in out else None
        _capi.check(self._lib.dlb_eval_batch(self._handle, X.data_ptr(), N, ptr('logposterior'), ptr('loglikelihood'), ptr('logprior'),
                                             ptr('x_solved_plan_order'), ptr('status'), ctypes.c_void_p(stream)))
        if own and 'x_solved_plan_order' in out:
            out['x_solved'] = out['x_solved_plan_order'] if self._identity_perm else out['x_solved_plan_order'][:, torch.as_tensor(self._unperm, device=X.device)]
        return out

    def eval_gather(self, X, gather, buf, slice_start, total, with_solved=False, out=None, wait=None):
        """Evaluate this rank's slice ``X[n, d]`` (rows ``slice_start : slice_start + n`` of a batch of ``total`` points) and let the finish
        kernel store logposterior (and x_solved, plan order) straight into buffer ``buf`` of EVERY rank's gather window
        (``distributed.PeerGather``; replaces reference base.py:335 / samplers/base.py:196-200).  ``X``: float64 CUDA tensor (enqueued on the
        current torch stream) or (pinned) host ndarray (staged by the library).  Optional local outputs in ``out``: ``loglikelihood``,
        ``logprior`` (float64), ``status`` (int32) of length n and ``x_solved_plan_order`` [n, ns] (this rank's own solved values), CUDA tensors or host
        ndarrays (host buffers are filled by the pipelined copy-out of the sub-batches).  ``wait``: buffer index whose acquire
        (``PeerGather.wait``) is enqueued behind this evaluation in the same call (pipelined loops: one library call per step)."""
        import torch
        out = out or {}
        if _is_torch_cuda(X):
            if X.dtype != torch.float64 or not X.is_contiguous() or X.dim() != 2 or X.shape[1] != self.d:
                raise ValueError('expected a contiguous float64 CUDA tensor of shape (n, {:d})'.format(self.d))
            xptr, device = (X.data_ptr() if X.shape[0] else None), X.device
        else:
            X = np.ascontiguousarray(X, dtype=np.float64)
            if X.ndim != 2 or X.shape[1] != self.d:
                raise ValueError('expected params of shape (n, {:d}), got {}'.format(self.d, X.shape))
            xptr, device = (X.ctypes.data if X.shape[0] else None), gather.device
        stream = torch.cuda.current_stream(device).cuda_stream

        def ptr(key):
            if key not in out: return None
            return out[key].data_ptr() if _is_torch_cuda(out[key]) else out[key].ctypes.data
        _capi.check(self._lib.dlb_eval_batch_gather(self._handle, xptr, X.shape[0], gather._handle, int(buf), int(slice_start),
                                                    int(total), int(bool(with_solved)), ptr('loglikelihood'), ptr('logprior'),
                                                    ptr('x_solved_plan_order') if self.ns else None, ptr('status'),
                                                    -1 if wait is None else int(wait), ctypes.c_void_p(stream)))

    def theory(self, X):
        """Flat theory vectors [N, n] at the solved parameters' defaults (reference ``likelihood.flattheory``)."""
        if self.lowered.split_dense: return self._theory_companion().theory(X)
        if _is_torch_cuda(X):
            import torch
            out = torch.empty((X.shape[0], self.ntheory), dtype=torch.float64, device=X.device)
            stream = torch.cuda.current_stream(X.device).cuda_stream
            _capi.check(self._lib.dlb_theory_batch(self._handle, X.data_ptr(), X.shape[0], out.data_ptr(), ctypes.c_void_p(stream)))
            return out
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.empty((X.shape[0], self.ntheory))
        _capi.check(self._lib.dlb_theory_batch(self._handle, X.ctypes.data, X.shape[0], out.ctypes.data, None))
        return out

    def features(self, gen, X):
        """Per-point features of generator ``gen`` [N, nfeat] (parity entry for the theory kernels)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        nfeat = int(self.lowered.gens[gen].nfeat)
        out = np.empty((X.shape[0], nfeat))
        _capi.check(self._lib.dlb_features_batch(self._handle, gen, X.ctypes.data, X.shape[0], out.ctypes.data, None))
        return out

    # ---- instrumentation ---------------------------------------------------------------------------------------------
    def executed_flops(self):
        """FP64 tensor-pipe FLOPs per evaluation executed by the contraction launch(es) (tile padding included, skipped tiles excluded)."""
        return float(self._lib.dlb_plan_executed_flops(self._handle))

    def set_timing(self, enabled=True):
        _capi.check(self._lib.dlb_set_timing(self._handle, int(enabled)))

    def last_kernel_ms(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        _capi.check(self._lib.dlb_last_kernel_ms(self._handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        f = ctypes.c_double()
        _capi.check(self._lib.dlb_last_fftlog_ms(ctypes.byref(f)))
        return dict(features=a.value, contract=b.value, solve=c.value, fftlog=f.value)


def premarginalize_precision(spec):
    """
    '.prec' pre-marginalisation at plan time, what the reference does once in ``BaseLikelihood._marginalize_precision``
    (desilike/likelihoods/base.py:257-312): with J = d flattheory / d x over the '.prec' parameters at the fiducial point
    (spec['fiducial']: the sampled parameters' default values) and their Gaussian priors N(loc, scale),

        precision <- P - (J P)^T (J P J^T + diag(1 / scale^2))^-1 (J P),   flatdata <- flatdata - sum_i loc_i J_i

    per precision block; the parameters then stay in the theory at their default values and are no longer solved.
    The Jacobian rows are evaluated ON THE GPU through ``dlb_theory_batch``: a temporary plan treats the '.prec' parameters as
    sampled columns, and J_i = theory(fiducial, x = e_i) - theory(fiducial, x = 0) is exact because the theory is linear in them.
    """
    solved = list(spec.get('solved', []))
    prec = [s for s in solved if s.get('mode') == 'prec']
    if not prec:
        return spec
    if 'fiducial' not in spec:
        raise ValueError("'.prec' parameters need spec['fiducial'] (default values of the sampled parameters)")
    names = list(spec['params'])
    pnames = [s['name'] for s in prec]
    # temporary plan: '.prec' parameters become sampled columns (factors of their terms)
    tmp = dict(spec)
    tmp['params'] = names + pnames
    tmp['priors'] = dict(spec.get('priors', {}))
    tmp['solved'] = [dict(s) for s in solved if s.get('mode') != 'prec']
    tmp['observables'] = []
    for obs in spec['observables']:
        o = dict(obs)
        o['terms'] = [dict(t, solved=None, factors=list(t.get('factors', [])) + [(t['solved'], 1)]) if t.get('solved') in pnames else t for t in obs['terms']]
        tmp['observables'].append(o)
    tmp.pop('fiducial', None)
    plan = Plan(tmp, with_theory=True)
    X = np.tile([float(spec['fiducial'][n]) for n in names] + [0.] * len(pnames), (len(pnames) + 1, 1))
    for i in range(len(pnames)): X[i + 1, len(names) + i] = 1.
    theory = plan.theory(X)
    plan.close()
    J = theory[1:] - theory[0]
    loc = np.array([float((s.get('prior') or {}).get('loc', 0.)) for s in prec])
    scale = np.array([float((s.get('prior') or {}).get('scale', np.inf)) if (s.get('prior') or {}).get('dist') == 'norm' else np.inf for s in prec])
    prior_prec = np.where(np.isinf(scale), 0., 1. / scale**2)
    precision = spec['precision']
    blocks = [np.asarray(p, dtype=np.float64) for p in precision] if isinstance(precision, (list, tuple)) else [np.asarray(precision, dtype=np.float64)]
    bsizes = [P.shape[0] for P in blocks]   # a block covers one observable or a run of consecutive observables
    starts = np.concatenate([[0], np.cumsum(bsizes)])
    # The reference updates the precision per LIKELIHOOD, with that likelihood's full precision (likelihoods/base.py:298-308): blocks that belong
    # to one likelihood (an ObservablesGaussianLikelihood built from per-observable covariances; spec['prec_groups']) are updated JOINTLY -- a
    # '.prec' parameter entering two of its observables couples them, the group becomes one dense block -- and a '.prec' parameter shared between
    # likelihoods is an error there too (:285-288).
    groups = [list(g) for g in spec.get('prec_groups', [])] or [[b] for b in range(len(blocks))]
    if sorted(b for g in groups for b in g) != list(range(len(blocks))) or any(g != list(range(g[0], g[0] + len(g))) for g in groups):
        raise ValueError("spec['prec_groups'] must partition the precision blocks into runs of consecutive blocks")
    member = [[i for i in range(len(prec)) if any(np.any(J[i, starts[b]:starts[b + 1]] != 0.) for b in g)] for g in groups]
    shared = [pnames[i] for i in range(len(prec)) if sum(i in m for m in member) > 1]
    if shared:
        raise ValueError('cannot use .prec for parameters that are shared between likelihoods (we would need to create a joint covariance matrix!); '
                         'common block is {}'.format(shared))
    new_prec = []
    for g, idx in zip(groups, member):
        sl = slice(starts[g[0]], starts[g[-1] + 1])
        touching = [b for b in g if any(np.any(J[i, starts[b]:starts[b + 1]] != 0.) for i in idx)]
        if not idx:
            new_prec += [blocks[b] for b in g]; continue
        if len(touching) <= 1 and len(g) > 1:   # the update stays inside one block of the group: keep the block structure
            for b in g:
                P = blocks[b]
                if b not in touching: new_prec.append(P); continue
                Jb = J[idx][:, starts[b]:starts[b + 1]]
                derivp = Jb * P if P.ndim == 1 else Jb.dot(P)
                H = derivp.dot(Jb.T) + np.diag(prior_prec[idx])
                new_prec.append((np.diag(P) if P.ndim == 1 else P) - derivp.T.dot(np.linalg.solve(H, derivp)))
            continue
        n = sl.stop - sl.start
        P = np.zeros((n, n))
        for b in g:
            o = starts[b] - sl.start
            P[o:o + bsizes[b], o:o + bsizes[b]] = np.diag(blocks[b]) if blocks[b].ndim == 1 else blocks[b]
        Jb = J[idx][:, sl]
        derivp = Jb.dot(P)
        H = derivp.dot(Jb.T) + np.diag(prior_prec[idx])
        new_prec.append(P - derivp.T.dot(np.linalg.solve(H, derivp)))
    shift = J.T.dot(loc)
    out = dict(spec)
    out['solved'] = [dict(s) for s in solved if s.get('mode') != 'prec']
    out['precision'] = new_prec if isinstance(precision, (list, tuple)) else new_prec[0]
    out.pop('prec_groups', None)
    values = {s['name']: float(s.get('value', 0.)) for s in prec}
    out['observables'], start = [], 0
    for obs in spec['observables']:
        o = dict(obs)
        n = len(obs['data'])
        o['data'] = np.asarray(obs['data'], dtype=np.float64) - shift[start:start + n]
        start += n
        terms = []
        for t in obs['terms']:
            if t.get('solved') in values:
                if values[t['solved']] == 0.: continue
                t = dict(t, solved=None, scale=float(t.get('scale', 1.)) * values[t['solved']])
            terms.append(t)
        o['terms'] = terms
        out['observables'].append(o)
    return out


def sample_solved(x_solved, hessian, rng=None):
    """Draw the analytically marginalised parameters around their best fit, x ~ N(x_best, (-H)^-1), per point -- what the reference's
    ``Chain.sample_solved`` does with the stored Hessians (desilike/samples/chain.py:46-72, :229-260).  ``x_solved`` [N, ns],
    ``hessian`` [N, ns, ns] as returned by ``Plan.eval(..., return_hessian=True)``."""
    rng = rng or np.random.default_rng()
    x_solved, hessian = np.asarray(x_solved, dtype=np.float64), np.asarray(hessian, dtype=np.float64)
    L = np.linalg.cholesky(-hessian)                       # -H = L L^T  ->  covariance = L^-T L^-1
    z = rng.standard_normal(x_solved.shape)
    return x_solved + np.linalg.solve(np.swapaxes(L, -1, -2), z[..., None])[..., 0]


def pinned_empty(shape, dtype=np.float64):
    """Page-locked host array (through torch's caching host allocator) for the host-buffer entry points."""
    import torch
    tdtype = {np.dtype('f8'): torch.float64, np.dtype('i4'): torch.int32}[np.dtype(dtype)]
    return torch.empty(shape, dtype=tdtype, pin_memory=True).numpy()


def launch_count():
    return int(_capi.load().dlb_launch_count())


def measure_fp64_peak(kind='dmma'):
    val = ctypes.c_double()
    _capi.check(_capi.load().dlb_measure_fp64_peak(kind if isinstance(kind, int) else (0 if kind == 'dmma' else 1), ctypes.byref(val)))
    return val.value
