"""Plan-time lowering: likelihood spec -> flat constant tables of ``dlb_plan_desc`` (include/desilike_b200.h).

All parameter-independent linear algebra of the chain is folded here, once, in FP64 on the host:

  * the precision matrix P (dense / block-diagonal / diagonal; built by the reference at
    desilike/likelihoods/base.py:605-631) is factorised P = L^T L and L is multiplied into
  * the window matrix W (reference desilike/observables/galaxy_clustering/window.py:459-468), the constant shot-noise /
    offset vectors (window.py:445-457), the bias-combine column weights (reference full_shape.py:545-550, :628-634) and
  * the emulator's linear output map (Taylor ``derivatives`` [T, ny], reference desilike/emulators/__init__.py:471-507;
    last MLP layer + inverse y-operations, reference desilike/emulators/conversion.py:18-33,76-84),

so that per point the GPU only evaluates features f(theta), bias monomials c(theta) and the contraction
``r = sum_t c_t f_t . M_t - r0`` (chi2 = |r|^2), plus the whitened Jacobian rows of the solved parameters.
Nothing here runs per point.

Likelihood spec (plain dicts / ndarrays; the same structure the oracle evaluates un-collapsed):
  spec = dict(params=[names], derived=[dict(name, expr)]  (optional: derived-expression parameters, '{b}**2'; terms / theories may use their names),
              priors={name: dict(dist, limits, loc, scale) | None},
              solved=[dict(name, value, prior, marg)], observables=[obs, ...],
              precision=[P_b per observable] | P joint [n, n] | diag [n])
  obs  = dict(theory=dict(kind='taylor'|'mlp'|'kaiser'|'bao_xi'|'bao_pk', params=[...], ...constants...),
              terms=[dict(scale, factors=[(param, pow)], solved=name|None, table=name|None, colvec=[nin]|None, outvec=[nobs]|None)],
              nin=int, wmatrix=[nobs, nin]|None, offset=[nobs]|None, data=[nobs], transform=None)
"""
import ctypes

import numpy as np

from . import _capi
from ._capi import f64, i32, as_ptr


class LoweringError(ValueError):
    pass


def _prior_tables(names, priors):
    n = len(names)
    kind, lo, hi = np.zeros(n, dtype=np.int32), np.full(n, -np.inf), np.full(n, np.inf)
    loc, scale = np.zeros(n), np.ones(n)
    for i, name in enumerate(names):
        prior = priors.get(name) if priors else None
        if prior is None: continue
        dist = prior.get('dist', 'uniform')
        limits = prior.get('limits', (-np.inf, np.inf))
        lo[i] = -np.inf if limits[0] is None else limits[0]
        hi[i] = np.inf if limits[1] is None else limits[1]
        if dist == 'uniform': kind[i] = 1
        elif dist == 'norm':
            kind[i] = 2; loc[i] = prior.get('loc', 0.); scale[i] = prior.get('scale', 1.)
        else:
            raise LoweringError('prior distribution {!r} of {} is not on the CUDA path (uniform / norm only, as the reference fast paths '
                                'parameter.py:2005-2009)'.format(dist, name))
    return kind, lo, hi, loc, scale


def _whitening(precision):
    """P = L^T L  ->  L (upper triangular for dense P, diagonal for 1-D P)."""
    precision = np.asarray(precision, dtype=np.float64)
    if precision.ndim == 1:
        if np.any(precision < 0.): raise LoweringError('negative entries in the diagonal precision')
        return np.diag(np.sqrt(precision))
    try:
        return np.linalg.cholesky(precision).T
    except np.linalg.LinAlgError:
        # positive SEMI-definite precision (e.g. '.prec' pre-marginalisation over a parameter without prior projects its template out of P):
        # any L with L^T L = P will do -- take the symmetric eigen-decomposition
        lam, V = np.linalg.eigh(0.5 * (precision + precision.T))
        if lam[0] < -1e-10 * max(lam[-1], 0.): raise LoweringError('precision matrix is not positive semi-definite')
        return np.sqrt(np.clip(lam, 0., None))[:, None] * V.T


def _merge_kaiser_terms(observables, names):
    """Tracer combine folded into the Kaiser generator.  The sampled terms of an observable that multiply the plain Kaiser tables without a
    column weight -- b1^2 pk_dd + 2 b1 pk_dt + pk_tt (reference full_shape.py:545-550) -- share one term matrix (the window), so their sum is ONE
    contraction of the combined table  sum_j scale_j prod theta^pow table_j,  which the Kaiser kernel forms directly in its mu quadrature (weight
    (b1 + f mu^2)^2): a third of the contraction FLOPs and a quarter of the feature columns.  Returns (observables with rewritten terms,
    {id(theory): segments}) -- the segments list every direct Kaiser theory's generator must output (csrc/theory_kernels.cu: kaiser_kernel)."""
    from .theory_lowering import KAISER_TABLE_ID
    plain = ('pk_dd', 'pk_dt', 'pk_tt')
    segments, out = {}, []
    for obs in observables:
        theory = obs['theory']
        if theory.get('kind') != 'kaiser':
            out.append(obs); continue
        segs = segments.setdefault(id(theory), [])

        def seg_index(seg):
            for i, s_ in enumerate(segs):
                if s_['flag'] == seg['flag'] and s_['combs'] == seg['combs']: return i
            seg['name'] = seg.get('name') or 'combo{:d}'.format(len(segs))
            segs.append(seg)
            return len(segs) - 1

        merge = [t for t in obs['terms'] if t.get('table') in plain and t.get('solved') is None and t.get('colvec') is None and t.get('featvec') is None and t.get('outvec') is None]
        terms = []
        if len(merge) >= 2:
            combs = tuple(sorted((KAISER_TABLE_ID[t['table']], float(t.get('scale', 1.)), tuple(sorted((names.index(n), int(p)) for n, p in t.get('factors', []))))
                                 for t in merge))
            i = seg_index(dict(flag=0, combs=combs))
            terms.append(dict(scale=1., factors=[], table=segs[i]['name']))
        for t in obs['terms']:
            if len(merge) >= 2 and any(t is m for m in merge): continue
            if t.get('table') in plain + ('pk11_l0',):
                seg_index(dict(name=t['table'], flag=int(t['table'] == 'pk11_l0'), combs=((KAISER_TABLE_ID[t['table']], 1., ()),)))
            terms.append(t)
        out.append(dict(obs, terms=terms))
    for segs in segments.values():
        if len(segs) > 8: raise LoweringError('more than 8 output segments for one Kaiser theory')
    return out, segments


class _Generator(object):
    """Host description of one feature generator + where each theory table lives in its feature segment."""

    def __init__(self, theory, names, priors=None, segments=None):
        self.kind_name = theory['kind']
        self.params = list(theory['params'])
        defaults = theory.get('defaults', {})
        self.param_idx = i32([names.index(p) if p in names else -1 for p in self.params])
        self.tables = {}  # name -> (foff, K, D [K, nin] or None for identity)
        self._names = names
        if self.kind_name == 'taylor':
            self.kind = _capi.GEN_TAYLOR
            center, powers = f64(theory['center']), i32(theory['powers'])
            T = powers.shape[0]
            self.param_default = f64([defaults.get(p, center[j]) for j, p in enumerate(self.params)])
            self.idata, self.ddata, self.nfeat = powers.ravel(), center, T + 2
            for name, deriv in theory['tables'].items():
                self.tables[name] = (2, T, f64(deriv).reshape(T, -1))
        elif self.kind_name == 'mlp':
            from .mlp import lower_mlp
            lower_mlp(self, theory, defaults)
        elif self.kind_name in ('kaiser', 'bao_xi', 'bao_pk'):
            from .theory_lowering import lower_theory, kaiser_set_segments
            lower_theory(self, theory, defaults, priors=priors)
            if segments and self.kind_name == 'kaiser': kaiser_set_segments(self, segments)
        else:
            raise LoweringError('unknown theory kind {!r}'.format(self.kind_name))

    def desc(self):
        g = _capi.GeneratorDesc()
        g.kind, g.nparams, g.nfeat = self.kind, len(self.params), int(self.nfeat)
        self.param_idx, self.param_default = i32(self.param_idx), f64(self.param_default)
        assert len(self.param_idx) == len(self.params) == len(self.param_default)
        self.idata, self.ddata = i32(self.idata), f64(self.ddata)
        g.param_idx, g.param_default = as_ptr(self.param_idx, ctypes.c_int32), as_ptr(self.param_default)
        g.ni, g.idata = int(self.idata.size), as_ptr(self.idata, ctypes.c_int32)
        g.nd, g.ddata = int(self.ddata.size), as_ptr(self.ddata)
        return g


def qform_tables(blk, force=False, coef_keys=None, const_features=None):
    """
    Quadratic-form evaluation of a likelihood block whose terms all contract the SAME K features f (one emulator per tracer: the usual case).

    The normal equations only need dot products of the slot vectors S_a = sum_{x in a} c_x(theta) f . M_x (minus r0 for a = 0) over the n whitened
    columns, and   (f . M_x) . (f . M_y) = f (M_x M_y^T) f^T   is a quadratic form in the K features.  With T terms the kernel then multiplies f by
    T (T + 1) / 2 matrices of K x K (+ a few linear columns) instead of T matrices of K x n: fewer columns whenever (T + 1) K / 2 < n -- e.g. 5 tracers x
    [66 features, 6 terms, n = 406]: 1336 columns instead of 2436 -- and every column is a single accumulator (no slots), so the faster single-slot
    kernel variants apply.  Exact algebra, done once here in FP64 like the whitening and window folding.

    Output columns, in n-tiles of 8 (one group per tile):
      * quadratic group of a (Gram entry, weight monomial c_x c_y): column k' holds sum over its term pairs of s_xy (M_x M_y^T)[:, k'], s_xy = 2 if x != y share
        a slot else 1; the epilogue multiplies column k' by f[k'] and adds the sum, weighted by c_x c_y, to the Gram entry pair(slot_x, slot_y).  Term pairs
        with the same entry and weight monomial (e.g. (1, b1^2) and (b1, b1)) share a group.  Features that are CONSTANT (``const_features``: the
        (1, 0) head of an MLP segment) get no quadratic column: a zero feature drops out, a constant one becomes a direct column -- for 64 hidden units
        the quadratic columns then fill exactly 8 tiles;
      * direct tiles: every column is added, with its own weight c_a c_b, to its own Gram entry: the constant-feature columns above, per term x
        -(2 if slot_x == 0 else 1) M_x r0 -> entry pair(0, slot_x) and M_x g_i -> entry of (constant row i, slot_x), weighted by c_x.
    ``const`` holds the theta-independent parts (|r0|^2, -r0 . g_i).  Returns None when not applicable (or not worthwhile, unless ``force``).
    """
    keys, white, R, n = blk['terms'], blk['white'], blk['nslots'], blk['ncols_white']
    T = len(keys)
    gi, foff, K = keys[0][2:5]
    if any(tuple(k[2:5]) != (gi, foff, K) for k in keys) or blk.get('entry_kind') is not None: return None
    g = np.asarray(blk['gconst_white'], dtype=np.float64).reshape(-1, n)
    nconst = g.shape[0]
    if T > 16: return None
    # quadratic columns: the non-constant features when they form one run starting at an even feature index (16-byte feature loads in the epilogue)
    consts = {k - foff: float(v) for k, v in (const_features or {}).items() if foff <= k < foff + K}
    kq = [k for k in range(K) if k not in consts]
    if not consts or not kq or kq != list(range(kq[0], kq[-1] + 1)) or (foff + kq[0]) % 2: consts, kq = {}, list(range(K))
    Kp = (len(kq) + 7) // 8 * 8
    if not force and (R < 2 or K > 128 or (T * (T + 1) // 2 * Kp + T * 8) >= 0.85 * T * (n + nconst)): return None
    r0 = np.asarray(blk['r0_white'], dtype=np.float64)
    nS = R * (R + 1) // 2

    def pair(a, b):
        a, b = min(a, b), max(a, b)
        return a * R - (a * (a - 1)) // 2 + (b - a)

    # term pairs whose Gram entry AND weight monomial c_x c_y coincide (e.g. (1, b1^2) and (b1, b1) in slot 0) share one group: their matrices add up
    inv_keys = {j: key for key, j in (coef_keys or {}).items()}

    def monomial(cx, cy):
        if coef_keys is None: return ('pair', cx, cy)   # no algebra available: never merge
        powers = {}
        for c in (cx, cy):
            for idx, pw in (inv_keys[c] if c >= 0 else ()): powers[idx] = powers.get(idx, 0) + pw
        return tuple(sorted((i, pw) for i, pw in powers.items() if pw != 0))

    merged = {}   # (entry, weight monomial) -> [matrix, coef x, coef y]
    for x in range(T):
        for y in range(x, T):
            a, b = keys[x][0], keys[y][0]
            C = (2. if (x != y and a == b) else 1.) * (white[x] @ white[y].T)   # column k' pairs with f[k']: sum_k f_k C[k, k'] f_k'
            key = (pair(a, b), monomial(keys[x][1], keys[y][1]))
            if key in merged: merged[key][0] = merged[key][0] + C
            else: merged[key] = [C, keys[x][1], keys[y][1]]
    # With constant features c (values v) the contraction only runs over the non-constant ones N (rows kq): for a matrix C
    #     f C f^T = f_N C_NN f_N^T + f_N . (C_Nc + C_cN^T) v + v^T C_cc v,
    # i.e. quadratic columns, one direct (linear) column and a constant; a linear form f . h = f_N . h_N + v . h_c likewise.  The constants are
    # theta-dependent only through their weight c_a c_b: ``wconst`` = (Gram entry, coef a, coef b, value), added once per point by the kernel.
    kc_idx, kc_val = list(consts), np.array([consts[k] for k in consts])
    direct, wconst = [], []   # direct: (matrix column over the rows kq, Gram entry, coef a, coef b)

    def linear(h, entry, ca, cb):   # f . h for a vector over ALL K features
        direct.append((h[kq], entry, ca, cb))
        if kc_idx and np.any(h[kc_idx] != 0.): wconst.append((float(kc_val.dot(h[kc_idx])), entry, ca, cb))

    for (entry, _), (C, cx, cy) in merged.items():
        if not kc_idx: continue
        lin = np.zeros(K)
        lin[kq] = (C[np.ix_(kq, kc_idx)] + C[np.ix_(kc_idx, kq)].T) @ kc_val
        if np.any(lin != 0.): direct.append((lin[kq], entry, cx, cy))
        cc = float(kc_val @ C[np.ix_(kc_idx, kc_idx)] @ kc_val)
        if cc != 0.: wconst.append((cc, entry, cx, cy))
    for x in range(T):
        a = keys[x][0]
        linear(-(2. if a == 0 else 1.) * (white[x] @ r0), pair(0, a), keys[x][1], -1)
        for i in range(nconst): linear(white[x] @ g[i], nS + i * R + a, keys[x][1], -1)
    # direct columns / weighted constants with the same Gram entry and weight monomial add up (e.g. -2 M_x r0 joins the constant-feature column of
    # the pair (x, '1' term)); then tiles of 8 columns with DISTINCT Gram entries each (the lanes of a tile update their entries concurrently)
    dmerged, wmerged = {}, {}
    for vec, entry, ca, cb in direct:
        key = (entry, monomial(ca, cb))
        if key in dmerged: dmerged[key][0] = dmerged[key][0] + vec
        else: dmerged[key] = [vec, entry, ca, cb]
    for val, entry, ca, cb in wconst:
        key = (entry, monomial(ca, cb))
        if key in wmerged: wmerged[key][0] += val
        else: wmerged[key] = [val, entry, ca, cb]
    dtiles = []
    for vec, entry, ca, cb in dmerged.values():
        for tile in dtiles:
            if len(tile) < 8 and all(e != entry for _, e, _, _ in tile): tile.append((vec, entry, ca, cb)); break
        else: dtiles.append([(vec, entry, ca, cb)])
    ndt = len(dtiles)
    ncolq = len(merged) * Kp + ndt * 8
    mat = np.zeros((len(kq), ncolq))
    tiles, groups, dent = [], [], []   # tiles: (group, first feature of the tile relative to the generator segment, valid columns)
    col = 0
    for (entry, _), (C, cx, cy) in merged.items():
        mat[:, col:col + len(kq)] = C[np.ix_(kq, kq)]
        groups.append((0, entry, cx, cy)); dent.append([[-1, -1, -1]] * 8)
        for j in range(Kp // 8): tiles.append((len(groups) - 1, foff + kq[0] + 8 * j, min(8, len(kq) - 8 * j)))
        col += Kp
    for t in range(ndt):
        ent = [[-1, -1, -1] for _ in range(8)]
        for j, (vec, entry, ca, cb) in enumerate(dtiles[t]):
            mat[:, col + j] = vec; ent[j] = [entry, ca, cb]
        groups.append((1, 0, -1, -1)); dent.append(ent)
        tiles.append((len(groups) - 1, foff + kq[0], 8))
        col += 8
    const = np.zeros(nS + R * nconst)
    const[pair(0, 0)] = r0.dot(r0)
    for i in range(nconst): const[nS + i * R] = -r0.dot(g[i])
    wlist = sorted(wmerged.values(), key=lambda w: w[1])   # by Gram entry
    return dict(matrix=mat, ncolq=ncolq, gen=gi, foff=foff + kq[0], K=len(kq), tiles=i32(tiles), groups=i32(groups), dent=i32(dent), const=f64(const),
                wconst_val=f64([w[0] for w in wlist]), wconst_idx=i32([[w[1], w[2], w[3]] for w in wlist]).reshape(-1, 3))


class Lowered(object):
    """Flat arrays + ctypes descriptor (keeps every buffer alive until the plan has been created)."""

    def __init__(self, spec, with_plain=True, reduce_rank=True, merge_terms=True, mlp_slices=0, qform='auto'):
        self.spec = spec
        self.mlp_slices = int(mlp_slices)
        self.qform = qform   # quadratic-form evaluation of blocks with solved parameters: 'auto' (when it multiplies fewer columns), True, False
        # derived-expression parameters (reference parameter.py:795-808, base.py:537): columns appended to every parameter row by a small
        # kernel; every index table below addresses the extended row, the batch itself has self.d_in columns
        from .expressions import compile_expression, ExpressionError
        self.input_names = list(spec['params'])
        self.d_in = len(self.input_names)
        names = self.names = self.input_names + [dp['name'] for dp in spec.get('derived', [])]
        if len(set(names)) != len(names): raise LoweringError('duplicate parameter names (sampled / derived): {}'.format(names))
        self.derived_programs, self.derived_consts = [], []
        for i, dp in enumerate(spec.get('derived', [])):
            try: prog, consts = compile_expression(dp['expr'], {n: j for j, n in enumerate(names[:self.d_in + i])})
            except ExpressionError as exc: raise LoweringError(str(exc))
            prog = [(op, arg if op != 1 else self._const_index(consts[arg])) for op, arg in prog]
            self.derived_programs.append(prog)
        self.d = len(names)
        priors = dict(spec.get('priors', {}) or {})
        for dp in spec.get('derived', []): priors[dp['name']] = None   # the prior is over the sampled parameters (reference parameter.py:1889-1897)
        self.prior = _prior_tables(names, priors)
        # ---- solved parameters: '.marg' first (so log det(-H[marg, marg]) is a leading-block Cholesky product) ----
        solved = list(spec.get('solved', []))
        order = [i for i, s in enumerate(solved) if s.get('marg', True)] + [i for i, s in enumerate(solved) if not s.get('marg', True)]
        self.solved_order = np.array(order, dtype=int)          # plan index -> spec index
        self.solved_names = [solved[i]['name'] for i in order]
        self.ns, self.nmarg = len(solved), sum(1 for s in solved if s.get('marg', True))
        sindex = {name: i for i, name in enumerate(self.solved_names)}
        self.x0 = f64([solved[i].get('value', 0.) for i in order])
        skind, slo, shi, sloc, sscale = _prior_tables(self.solved_names, {s['name']: s.get('prior') for s in solved})
        sscale = np.where(skind == 2, sscale, np.inf)
        self.xprior = (skind, sloc, sscale, slo, shi)
        # ---- generators (one per distinct theory object) ----
        self.gens, gen_of = [], {}
        observables, segments = list(spec['observables']), {}
        if merge_terms: observables, segments = _merge_kaiser_terms(observables, names)
        for obs in observables:
            key = id(obs['theory'])
            if key not in gen_of:
                gen_of[key] = len(self.gens)
                self.gens.append(_Generator(obs['theory'], names, priors=spec.get('priors'), segments=segments.get(key)))
        # ---- coefficients ----
        self.coef_keys = {}

        def coef_index(factors):
            key = tuple(sorted((names.index(n), int(p)) for n, p in factors))
            if not key: return -1
            if len(key) > _capi.DLB_COEF_MAXF: raise LoweringError('too many factors in a bias monomial')
            if key not in self.coef_keys: self.coef_keys[key] = len(self.coef_keys)
            return self.coef_keys[key]

        # ---- blocks ----
        prec = spec['precision']
        if isinstance(prec, (list, tuple)):   # one block per run of consecutive observables (usually one observable each; a SumLikelihood
            groups, io = [], 0                # member with a joint covariance over several observables gives one block for all of them)
            for p in prec:
                p = np.asarray(p, dtype=np.float64)
                group, size = [], 0
                while size < p.shape[0] and io < len(observables):
                    group.append(observables[io]); size += len(observables[io]['data']); io += 1
                if size != p.shape[0]: raise LoweringError('precision block of size {:d} does not match a run of observables'.format(p.shape[0]))
                groups.append((group, p))
            if io != len(observables): raise LoweringError('precision blocks do not cover all observables')
        else:
            groups = [(list(observables), np.asarray(prec, dtype=np.float64))]
        self.blocks = []
        self.split_dense = False   # some dense precision block was split over several passes (theory vectors then come from a companion plan)
        self.m0_extra = None
        self.two_stage = False   # any transformed observable: theory -> transform -> whitening as two contraction launches
        for group, P in groups:
            ncols = sum(len(o['data']) for o in group)
            if P.shape[0] != ncols: raise LoweringError('precision size {} != data size {}'.format(P.shape, ncols))
            L = _whitening(P)
            terms = {}   # (slot, coef, gen, foff, K) -> plain matrix [K, ncols]
            slot_of, const_rows = {}, {}
            const_plain = np.zeros(ncols)
            data = np.concatenate([f64(o['data']) for o in group])
            c0 = 0
            tdata = np.zeros(ncols)   # flat data of the 'cubic'-transformed columns (reference power_spectrum.py:400-404), 0 elsewhere
            for obs in group:
                n_o, nin = len(obs['data']), int(obs['nin'])
                if obs.get('transform') is not None:
                    if obs['transform'] != 'cubic': raise LoweringError('unknown observable transform {!r}'.format(obs['transform']))
                    if np.any(f64(obs['data']) == 0.): raise LoweringError("transform 'cubic' divides by the data vector: zero entries found")
                    tdata[c0:c0 + n_o] = f64(obs['data'])
                    self.two_stage = True
                W = f64(obs['wmatrix']) if obs.get('wmatrix') is not None else None
                if W is None and nin != n_o: raise LoweringError('wmatrix=None requires nin == len(data)')
                gi = gen_of[id(obs['theory'])]
                gen = self.gens[gi]
                if obs.get('offset') is not None: const_plain[c0:c0 + n_o] += f64(obs['offset'])

                def add(slot, coef, foff, K, mat_o):
                    key = (slot, coef, gi, foff, K)
                    if key not in terms: terms[key] = np.zeros((K, ncols))
                    terms[key][:, c0:c0 + n_o] += mat_o

                for term in obs['terms']:
                    scale = float(term.get('scale', 1.))
                    factors = list(term.get('factors', []))
                    colvec = f64(term['colvec']) if term.get('colvec') is not None else None
                    # featvec: weights in the FEATURE space of the table, applied before its fold matrix (counter terms (k / kp)^pow pk11 of a
                    # correlation-function theory: the weight acts on P(kin), the P -> xi operator comes after); colvec acts on the theory vector
                    featvec = f64(term['featvec']) if term.get('featvec') is not None else None
                    sname = term.get('solved')
                    if term.get('table') is not None:
                        foff, K, D = gen.tables[term['table']]
                        if featvec is not None and featvec.shape != (K,): raise LoweringError('featvec must have the length of its table ({:d})'.format(K))
                        if D is None:   # identity features: the table itself is the feature sub-segment
                            cv = np.ones(nin) if colvec is None else colvec
                            if featvec is not None: cv = cv * featvec
                            mat = scale * (cv[:, None] * (W.T if W is not None else np.eye(nin)))
                        else:
                            Dc = D if featvec is None else featvec[:, None] * D
                            if colvec is not None: Dc = Dc * colvec
                            mat = scale * (Dc @ W.T if W is not None else Dc)
                        theta_dep = True
                    else:
                        if term.get('outvec') is not None:   # output-space template, added after the window (systematic templates, reference window.py:470-473)
                            vec = scale * f64(term['outvec'])
                            if vec.shape != (n_o,): raise LoweringError('outvec must have the length of the observable ({:d})'.format(n_o))
                        else:
                            cv = np.ones(nin) if colvec is None else colvec
                            vec = scale * (W @ cv if W is not None else cv)
                        foff, K, mat = 0, 1, vec[None, :]
                        theta_dep = bool(factors)
                    if sname is None:
                        if theta_dep: add(0, coef_index(factors), foff, K, mat)
                        else: const_plain[c0:c0 + n_o] += mat[0]
                        continue
                    si = sindex[sname]
                    if theta_dep:
                        if si not in slot_of: slot_of[si] = 1 + len(slot_of)
                        add(slot_of[si], coef_index(factors), foff, K, mat)
                        if self.x0[si] != 0.: add(0, coef_index(factors), foff, K, self.x0[si] * mat)
                    else:
                        if si in slot_of: raise LoweringError('solved parameter {} mixes constant and theta-dependent templates'.format(sname))
                        row = const_rows.setdefault(si, np.zeros(ncols))
                        row[c0:c0 + n_o] += mat[0]
                        if self.x0[si] != 0.: const_plain[c0:c0 + n_o] += self.x0[si] * mat[0]
                c0 += n_o
            for si in const_rows:
                if si in slot_of: raise LoweringError('solved parameter {} mixes constant and theta-dependent templates'.format(self.solved_names[si]))
            if np.any(tdata != 0.) and const_rows:
                # transformed observable: the Jacobian rows are u'(t(theta)) dt/dx_i (the reference differentiates flatdiff through the transform,
                # likelihoods/base.py:152-158), so constant templates become slots as well -- riding on the constant-one feature of the segment
                gi0 = gen_of[id(group[0]['theory'])]
                for si in sorted(const_rows):
                    slot_of[si] = 1 + len(slot_of)
                    terms[(slot_of[si], -1, gi0, 0, 1)] = const_rows[si][None, :].copy()
                const_rows = {}
            nslots = 1 + len(slot_of)
            if nslots > 8:
                # dense joint precision coupling more theta-dependent solved templates than one pass holds (8 accumulator slots): split
                if self.two_stage: raise LoweringError('observable transform together with a split dense precision is not supported')
                self.blocks += self._lower_dense_split(group, P, terms, slot_of, const_rows, const_plain, data)
                self.split_dense = True
                continue
            blk = dict(ncols=ncols, nslots=nslots)
            keys = sorted(terms.keys())
            blk['terms'] = keys
            blk['plain'] = [terms[k] for k in keys]
            blk['white'] = [terms[k] @ L.T for k in keys]
            blk['r0_white'] = L @ (data - const_plain)
            blk['ncols_white'] = ncols
            Ktot = sum(k[4] for k in keys)
            if reduce_rank and nslots == 1 and not const_rows and ncols > Ktot + 1:
                # Rank reduction of the whitened residual.  r = g . M - r0 with g = (c_t f_t)_t in R^Ktot lives in the row space of the
                # stacked term matrices M [Ktot, n]; with the thin QR factorisation M^T = Q R (Q [n, Ktot] orthonormal)
                #     |g M - r0|^2 = |g R^T - Q^T r0|^2 + |r0 - Q Q^T r0|^2,
                # i.e. the same chi2 from Ktot + 1 columns instead of n (the last column carries the constant |r0_perp|), and R^T is
                # lower triangular.  Exact linear algebra, done once here in FP64 -- like the whitening and window folding above.
                M = np.concatenate(blk['white'], axis=0)
                Q, R = np.linalg.qr(M.T)
                proj = Q.T @ blk['r0_white']
                perp = blk['r0_white'] - Q @ proj
                Rt = np.concatenate([R.T, np.zeros((Ktot, 1))], axis=1)
                off, white = 0, []
                for k in keys:
                    white.append(np.ascontiguousarray(Rt[off:off + k[4]])); off += k[4]
                blk['white'], blk['r0_white'], blk['ncols_white'] = white, np.append(proj, np.sqrt(perp.dot(perp))), Ktot + 1
            blk['t0_plain'] = const_plain
            blk['transform_data'], blk['whitening_t'], blk['r0_data_white'] = tdata, np.ascontiguousarray(L.T), L @ data
            inv = sorted(slot_of.items(), key=lambda kv: kv[1])
            blk['slot_solved'] = i32([si for si, _ in inv])
            blk['const_solved'] = i32(sorted(const_rows.keys()))
            blk['gconst_white'] = f64([L @ const_rows[si] for si in sorted(const_rows.keys())]).reshape(len(const_rows), ncols)
            if self.qform is not False and not self.two_stage and blk['ncols_white'] == ncols:
                blk['qform'] = qform_tables(blk, force=self.qform is True, coef_keys=self.coef_keys,
                                            const_features=getattr(self.gens[keys[0][2]], 'const_features', None))
            self.blocks.append(blk)
        if any(b.get('qform') is None for b in self.blocks):   # all blocks or none: one kernel variant per launch
            for b in self.blocks: b['qform'] = None
        self.ntheory = sum(len(o['data']) for o in observables)
        self.with_plain = with_plain
        self._build_desc()

    def _lower_dense_split(self, group, P, terms, slot_of, const_rows, const_plain, data):
        """
        Dense joint precision P over several observables whose theta-dependent solved templates (counter-term rows ~ pk11(theta)) number more than
        the 7 one contraction pass can hold next to the residual (SURVEY.md section 8d, C4 'dense 2000 x 2000' stress variant: 5 tracers x 3 rows).

        The Jacobian rows are observable-local BEFORE whitening, so instead of whitened rows J~ = L J the plan carries, per observable o,
            S-slots   S_r = (t - d)[cols_o],  S_i = J_i[cols_o]  for the rows i of o            (un-whitened, local)
            y-slots   y_j[cols_o] = (P S_j)[cols_o]  for j in {r} + all rows                     (precision-weighted, dense in j's observable)
        and every quantity of the normal equations (reference likelihoods/base.py:165-187) is a sum over observables of S . y products:
            chi2 = sum_o S_r . y_r,   gradient_i = -S_i . y_r (o = observable of i),   M_ij = S_i . y_j (once per unordered pair),
        constant rows g_c (stochastic terms) enter through the auxiliary columns w_c = (P g_c)[cols_o] of the S-slots and a plan-constant Gram
        g_c . P g_c' (dlb_plan_desc::m0_extra).  The y-slots are spread over passes of (8 - number of S-slots) each: one block per (observable, pass),
        its Gram partials tagged explicitly (dlb_block_desc::entry_kind).
        """
        from collections import OrderedDict
        sizes = [len(o['data']) for o in group]
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
        ncols = int(starts[-1])
        P = np.diag(P) if P.ndim == 1 else np.asarray(P, dtype=np.float64)
        theta_slots = sorted(slot_of.items(), key=lambda kv: kv[1])        # [(solved index, slot)]
        # observable of every theta-dependent row: its plain matrices must vanish outside one observable's columns
        obs_of = {}
        for key, mat in terms.items():
            slot = key[0]
            if slot == 0: continue
            touched = [io for io in range(len(group)) if np.any(mat[:, starts[io]:starts[io + 1]] != 0.)]
            if len(touched) > 1: raise LoweringError('split dense precision: a theta-dependent solved template spans several observables')
            if touched:
                if obs_of.setdefault(slot, touched[0]) != touched[0]: raise LoweringError('split dense precision: a solved template spans several observables')
        slot_solved = {slot: si for si, slot in theta_slots}
        yslots = [0] + [slot for _, slot in theta_slots]                    # global slot ids with a precision-weighted copy
        const_list = sorted(const_rows.keys())
        W = {si: P @ const_rows[si] for si in const_list}                   # w_c = P g_c
        m0 = np.zeros((self.ns, self.ns))
        for a in const_list:
            for b in const_list: m0[a, b] += const_rows[a].dot(W[b])
        self.m0_extra = m0 if self.m0_extra is None else self.m0_extra + m0
        r0_full = data - const_plain
        Pr0 = P @ r0_full
        blocks = []
        for io, obs in enumerate(group):
            c0, c1 = int(starts[io]), int(starts[io + 1])
            n_o = c1 - c0
            local = [0] + [slot for slot in sorted(obs_of) if obs_of[slot] == io]      # S-slots of this observable (global slot ids)
            if len(local) > 6: raise LoweringError('split dense precision: more than 5 theta-dependent solved templates in one observable')
            per_pass = 8 - len(local)
            npass = (len(yslots) + per_pass - 1) // per_pass
            gen_o = None
            for key in terms:
                if np.any(terms[key][:, c0:c1] != 0.): gen_o = key[2]; break
            cchunks = [const_list[q::npass] for q in range(npass)]              # constant rows spread over the passes (aux columns of the S-slots)
            for q in range(npass):
                ys = yslots[q * per_pass:(q + 1) * per_pass]
                nsl = len(local) + len(ys)
                bterms = OrderedDict()
                for key, mat in terms.items():
                    slot, coef, gi, foff, K = key
                    if slot in local:
                        sub = mat[:, c0:c1]
                        if np.any(sub != 0.): bterms[(local.index(slot), coef, gi, foff, K)] = bterms.get((local.index(slot), coef, gi, foff, K), 0.) + sub
                    if slot in ys:
                        sub = mat @ P[:, c0:c1]
                        k2 = (len(local) + ys.index(slot), coef, gi, foff, K)
                        bterms[k2] = bterms.get(k2, 0.) + sub
                if 0 in ys:   # y_r = P (t - d): the constant part -P (d - const) rides on the constant-one feature of the segment
                    k2 = (len(local) + ys.index(0), -1, gen_o, 0, 1)
                    bterms[k2] = bterms.get(k2, 0.) - Pr0[None, c0:c1]
                keys = sorted(bterms.keys())
                cs = cchunks[q]
                npart = nsl * (nsl + 1) // 2 + nsl * len(cs)
                kind, ei, ej = np.full(npart, 3, dtype=np.int32), np.zeros(npart, dtype=np.int32), np.zeros(npart, dtype=np.int32)

                def pair(a, b): return a * nsl - (a * (a - 1)) // 2 + (b - a)
                for ia, sa in enumerate(local):
                    for jb, sb in enumerate(ys):
                        e = pair(ia, len(local) + jb)
                        if sa == 0 and sb == 0: kind[e] = 0
                        elif sa != 0 and sb == 0: kind[e], ei[e] = 1, slot_solved[sa]            # gradient_i -= S_i . y_r
                        elif sa != 0 and sb != 0:
                            i, j = slot_solved[sa], slot_solved[sb]
                            # every unordered pair once: take it where i's observable precedes j's (or i <= j inside one observable)
                            if (obs_of[sa], i) <= (obs_of[sb], j): kind[e], ei[e], ej[e] = 2, i, j
                nS = nsl * (nsl + 1) // 2
                for ic, si in enumerate(cs):
                    for ia, sa in enumerate(local):
                        e = nS + ic * nsl + ia
                        if sa == 0: kind[e], ei[e] = 1, si                                        # gradient_c -= r . P g_c
                        else: kind[e], ei[e], ej[e] = 2, slot_solved[sa], si                      # M[i, c] += S_i . P g_c
                blk = dict(ncols=n_o, nslots=nsl, terms=keys, plain=[np.zeros((k[4], n_o)) for k in keys], white=[np.ascontiguousarray(bterms[k]) for k in keys],
                           r0_white=r0_full[c0:c1].copy(), ncols_white=n_o, t0_plain=np.zeros(n_o), transform_data=np.zeros(n_o),
                           whitening_t=np.zeros((1, 1)), r0_data_white=np.zeros(n_o), slot_solved=i32(np.zeros(nsl - 1)), const_solved=i32(cs),
                           gconst_white=f64([W[si][c0:c1] for si in cs]).reshape(len(cs), n_o), entry_kind=kind, entry_i=ei, entry_j=ej, skip_const_gram=1)
                blocks.append(blk)
        return blocks

    def _const_index(self, value):
        if value not in self.derived_consts: self.derived_consts.append(value)
        return self.derived_consts.index(value)

    def _build_desc(self):
        keep = self._keep = []
        desc = _capi.PlanDesc()
        desc.abi_version, desc.d = _capi.DLB_ABI_VERSION, self.d_in
        nder = len(self.derived_programs)
        if nder:
            off = i32(np.cumsum([0] + [len(p) for p in self.derived_programs]))
            code = i32(np.array([pair for p in self.derived_programs for pair in p]).ravel())
            consts = f64(self.derived_consts if self.derived_consts else [0.])
            keep += [off, code, consts]
            desc.nderived, desc.derived_off, desc.derived_code, desc.derived_const = nder, as_ptr(off, ctypes.c_int32), as_ptr(code, ctypes.c_int32), as_ptr(consts)
        kind, lo, hi, loc, scale = [np.ascontiguousarray(a) for a in self.prior]
        keep += [kind, lo, hi, loc, scale]
        desc.prior_kind, desc.prior_lo, desc.prior_hi = as_ptr(kind, ctypes.c_int32), as_ptr(lo), as_ptr(hi)
        desc.prior_loc, desc.prior_scale = as_ptr(loc), as_ptr(scale)
        gens = (_capi.GeneratorDesc * len(self.gens))(*[g.desc() for g in self.gens])
        desc.ngen, desc.gens = len(self.gens), gens
        ncoef = len(self.coef_keys)
        coefs = (_capi.CoefDesc * max(ncoef, 1))()
        for key, j in self.coef_keys.items():
            coefs[j].scale, coefs[j].nf = 1., len(key)
            for f, (idx, power) in enumerate(key):
                coefs[j].idx[f], coefs[j].pow[f] = idx, power
        desc.ncoef, desc.coefs = ncoef, coefs
        blocks = (_capi.BlockDesc * len(self.blocks))()
        for b, blk in zip(blocks, self.blocks):
            nt = len(blk['terms'])
            terms = (_capi.TermDesc * nt)()
            moff = moff_plain = 0
            for t, key, mat, mat_plain in zip(terms, blk['terms'], blk['white'], blk['plain']):
                t.slot, t.coef, t.gen, t.foff, t.K = key
                t.moff, t.moff_plain = moff, moff_plain
                moff += mat.size
                moff_plain += mat_plain.size
            white = f64(np.concatenate([m.ravel() for m in blk['white']]))
            plain = f64(np.concatenate([m.ravel() for m in blk['plain']])) if ((self.with_plain or self.two_stage) and not self.split_dense) else None
            r0, t0 = f64(blk['r0_white']), f64(blk['t0_plain'])
            ss, cs, gc = i32(blk['slot_solved']), i32(blk['const_solved']), f64(blk['gconst_white'])
            keep += [terms, white, plain, r0, t0, ss, cs, gc]
            if self.two_stage:
                td, wt, rd = f64(blk['transform_data']), f64(blk['whitening_t']), f64(blk['r0_data_white'])
                keep += [td, wt, rd]
                b.transform_data, b.whitening_t, b.r0_data_white = as_ptr(td), as_ptr(wt), as_ptr(rd)
            b.ncols, b.nslots, b.nterms, b.terms = blk['ncols'], blk['nslots'], nt, terms
            b.ncols_white = blk['ncols_white']
            b.matrices_white, b.matrices_plain = as_ptr(white), as_ptr(plain)
            b.r0_white, b.t0_plain = as_ptr(r0), as_ptr(t0)
            b.slot_solved, b.nconst, b.const_solved, b.gconst_white = as_ptr(ss, ctypes.c_int32), len(cs), as_ptr(cs, ctypes.c_int32), as_ptr(gc)
            if blk.get('qform') is not None:
                q = blk['qform']
                qm, qt, qg, qd, qc = f64(q['matrix']), i32(q['tiles']), i32(q['groups']), i32(q['dent']), f64(q['const'])
                qwv, qwi = f64(q['wconst_val']), i32(q['wconst_idx'])
                keep += [qm, qt, qg, qd, qc, qwv, qwi]
                b.q_ncols, b.q_ngroups = int(q['ncolq']), len(qg)
                b.q_foff, b.q_K, b.q_nw = int(q['foff']), int(q['K']), len(qwv)
                if len(qwv): b.q_wconst_val, b.q_wconst_idx = as_ptr(qwv), as_ptr(qwi, ctypes.c_int32)
                b.q_matrix, b.q_tiles, b.q_groups, b.q_dent, b.q_const = as_ptr(qm), as_ptr(qt, ctypes.c_int32), as_ptr(qg, ctypes.c_int32), as_ptr(qd, ctypes.c_int32), as_ptr(qc)
            if blk.get('entry_kind') is not None:
                ek, ei, ej = i32(blk['entry_kind']), i32(blk['entry_i']), i32(blk['entry_j'])
                keep += [ek, ei, ej]
                b.entry_kind, b.entry_i, b.entry_j = as_ptr(ek, ctypes.c_int32), as_ptr(ei, ctypes.c_int32), as_ptr(ej, ctypes.c_int32)
                b.skip_const_gram = int(blk.get('skip_const_gram', 0))
        desc.nblocks, desc.blocks = len(self.blocks), blocks
        skind, sloc, sscale, slo, shi = [np.ascontiguousarray(a) for a in self.xprior]
        keep += [gens, coefs, blocks, skind, sloc, sscale, slo, shi, self.x0]
        desc.ns, desc.nmarg = self.ns, self.nmarg
        desc.x0, desc.xprior_kind = as_ptr(self.x0), as_ptr(skind, ctypes.c_int32)
        desc.xprior_loc, desc.xprior_scale, desc.xprior_lo, desc.xprior_hi = as_ptr(sloc), as_ptr(sscale), as_ptr(slo), as_ptr(shi)
        desc.mlp_slices = self.mlp_slices
        if self.m0_extra is not None:
            m0 = f64(self.m0_extra)
            keep.append(m0)
            desc.m0_extra = as_ptr(m0)
        self.desc = desc

    def flops_per_point(self):
        """Algorithmic FLOPs per evaluation of the folded chain (2 per multiply-add), by stage; DESIGN.md section 'roofline'."""
        contract = 0
        gram = 0
        for blk in self.blocks:
            R, nc = blk['nslots'], len(blk['const_solved'])
            if blk.get('qform') is not None:   # quadratic-form evaluation: ONE matrix [K, ncolq]; the epilogue multiplies every column by its feature
                contract += 2 * blk['qform']['K'] * blk['qform']['ncolq']
                gram += 2 * blk['qform']['ncolq']
                continue
            for (slot, coef, gen, foff, K), mat in zip(blk['terms'], blk['white']):
                # + nc auxiliary columns per term: the dot products with the constant whitened Jacobian rows come out of the
                # contraction itself (csrc/plan.cu: build_block_core).  Rank-reduced blocks are lower triangular: only the non-zero
                # entries count (the kernel skips the leading all-zero tiles of every column block and executes slightly more)
                contract += 2 * (K * (blk['ncols_white'] + nc) if blk['ncols_white'] == blk['ncols'] else int(np.count_nonzero(mat)))
            gram += 2 * blk['ncols_white'] * (R * (R + 1) // 2)
            if self.two_stage: contract += 2 * blk['ncols']**2   # second stage: whitening of the transformed theory vector
        solve = self.ns**3 // 3 + 2 * self.ns**2
        return dict(contract=contract, gram=gram, solve=solve, total=contract + gram + solve)
