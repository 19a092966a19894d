"""Oracle evaluation of a likelihood *spec* (the un-collapsed chain, stage by stage, FP64 NumPy).

TEST INFRASTRUCTURE — the checker, never the product (see oracle/core.py header).

The spec (plain dicts / ndarrays; schema in DESIGN.md "likelihood spec") describes what the reference pipeline of
SURVEY.md section 3.1 computes for one point:

    tables(theta)                                (emulator predict | Kaiser | damped BAO -> FFTLog)
 -> P = sum_terms coef(theta, x) * colvec * table        (tracer combine, reference full_shape.py:545-550, :628-634; bao.py:527-529, :905)
 -> flat = W . P + offset                         (reference window.py:459-473)
 -> d = concat(flat) - data;  logL = -1/2 d^T P d (reference likelihoods/base.py:658-664, :13-17)
 -> analytic marginalisation over linear x        (reference likelihoods/base.py:129-200, :314-413)
 -> + prior(theta)                                (reference parameter.py:1889-1897, :1994-2017)
"""
import numpy as np

from . import core


def _coef(term, X, names, xvalues):
    """Per-point scalar coefficient of a term: scale * prod theta^pow * (solved value)."""
    c = np.full(X.shape[0], float(term.get('scale', 1.)))
    for name, power in term.get('factors', []):
        c = c * X[:, names.index(name)]**power
    if term.get('solved') is not None:
        c = c * xvalues[term['solved']]
    return c


def theory_tables(theory, X, names):
    """-> dict table name -> [N, nin]."""
    kind = theory['kind']
    cols = [names.index(p) for p in theory['params']]
    Xe = X[:, cols]
    if kind == 'taylor':
        mono = core.taylor_monomials(Xe, theory['center'], theory['powers'])
        return {name: mono @ deriv.reshape(deriv.shape[0], -1) for name, deriv in theory['tables'].items()}
    if kind == 'mlp':
        y = core.mlp_predict(Xe, theory['mlp'])
        return {name: (y[:, sl[0]] @ np.asarray(sl[1], dtype='f8') if isinstance(sl, tuple) else y[:, sl]) for name, sl in theory['tables'].items()}
    if kind == 'kaiser':
        tpl = theory['template']
        mu, wmu = core.legendre_weights(theory['ells'], theory['nmu'])
        # correlation function (reference full_shape.py:336-365, :553-574): the Kaiser tables on kin, each pushed through get_corr (base.py:127-136)
        p2c = core.PowerToCorr(theory['s'], theory['ells'], kin=theory.get('kin'), interp_order=int(theory.get('interp_order', 1))) if theory.get('s') is not None else None
        k = np.asarray(theory['k'], dtype='f8') if p2c is None else p2c.kin
        nx = k.size if p2c is None else len(theory['s'])
        out = {name: np.empty((X.shape[0], len(theory['ells']) * nx)) for name in ['pk_dd', 'pk_dt', 'pk_tt', 'pk11_l0']}
        for ip in range(X.shape[0]):
            p = dict(zip(theory['params'], Xe[ip]))
            p = {**theory.get('defaults', {}), **p}
            qpar, qper = core.ap_qpar_qper(tpl['apmode'], p, tpl.get('eta', 1. / 3.))
            pk_t = tpl['pk_dd_fid'] * core.shapefit_factor(tpl['k'], p.get('dm', 0.), p.get('dn', 0.), tpl['a'], tpl['kp'])
            f = tpl['f_fid'] * p.get('df', 1.)
            tabs = core.kaiser_tables(k, mu, wmu, tpl['k'], pk_t, f, qpar, qper, p.get('sigmapar', 0.), p.get('sigmaper', 0.), interp=theory.get('interp', tpl.get('interp', 'scipy')))
            pk11 = np.tile(tabs['pk_dd'][list(theory['ells']).index(0)], (len(theory['ells']), 1)) if 0 in theory['ells'] else 0. * tabs['pk_dd']
            for name, tab in [('pk_dd', tabs['pk_dd']), ('pk_dt', tabs['pk_dt']), ('pk_tt', tabs['pk_tt']), ('pk11_l0', pk11)]:
                out[name][ip] = (tab if p2c is None else p2c(tab)).ravel()
                if p2c is not None:   # the P(kin)-space tables too: terms with feature-space weights ('featvec') transform weight x table
                    out.setdefault(name + '@kin', np.empty((X.shape[0], tab.size)))[ip] = tab.ravel()
        if p2c is not None: out['@p2c'] = p2c
        return out
    if kind == 'bao_xi':
        tpl = theory['template']
        mu, wmu = core.legendre_weights(theory['ells'], theory['nmu'])
        p2c = core.PowerToCorr(theory['s'], theory['ells'], kin=theory.get('kin'), interp_order=int(theory.get('interp_order', 1)))
        out = {'corr': np.empty((X.shape[0], len(theory['ells']) * len(theory['s'])))}
        for ip in range(X.shape[0]):
            p = dict(zip(theory['params'], Xe[ip]))
            p = {**theory.get('defaults', {}), **p}
            qpar, qper = core.ap_qpar_qper(tpl['apmode'], p, tpl.get('eta', 1. / 3.))
            power = core.bao_damped_power(p2c.kin, mu, wmu, tpl['k'], tpl['pk_dd_fid'], tpl['pknow_dd_fid'], tpl['f_fid'] * p.get('df', 1.),
                                          qpar, qper, b1=p.get('b1', 1.), dbeta=p.get('dbeta', 1.), sigmas=p.get('sigmas', 0.),
                                          sigmapar=p.get('sigmapar', 9.), sigmaper=p.get('sigmaper', 6.), mode=theory.get('mode', ''),
                                          smoothing_radius=theory.get('smoothing_radius', 15.))
            out['corr'][ip] = p2c(power).ravel()
        return out
    if kind == 'bao_pk':
        tpl = theory['template']
        mu, wmu = core.legendre_weights(theory['ells'], theory['nmu'])
        k = np.asarray(theory['k'], dtype='f8')
        out = {'power': np.empty((X.shape[0], len(theory['ells']) * k.size))}
        for ip in range(X.shape[0]):
            p = dict(zip(theory['params'], Xe[ip]))
            p = {**theory.get('defaults', {}), **p}
            qpar, qper = core.ap_qpar_qper(tpl['apmode'], p, tpl.get('eta', 1. / 3.))
            out['power'][ip] = core.bao_damped_power(k, mu, wmu, tpl['k'], tpl['pk_dd_fid'], tpl['pknow_dd_fid'], tpl['f_fid'] * p.get('df', 1.),
                                                     qpar, qper, b1=p.get('b1', 1.), dbeta=p.get('dbeta', 1.), sigmas=p.get('sigmas', 0.),
                                                     sigmapar=p.get('sigmapar', 9.), sigmaper=p.get('sigmaper', 6.), mode=theory.get('mode', ''),
                                                     smoothing_radius=theory.get('smoothing_radius', 15.)).ravel()
        return out
    raise ValueError('unknown theory kind {}'.format(kind))


def observable_flat(obs, X, names, xvalues, tables=None, only_solved=None):
    """flat theory [N, nobs] of one observable.  ``only_solved``: return d flat / d x_name instead (Jacobian row)."""
    if tables is None:
        tables = theory_tables(obs['theory'], X, names)
    P = np.zeros((X.shape[0], obs['nin']), dtype='f8')
    out = 0.   # systematic templates: added after the window (reference observables/galaxy_clustering/window.py:470-473, :1285-1287)
    for term in obs['terms']:
        if only_solved is not None:
            if term.get('solved') != only_solved: continue
            c = _coef(term, X, names, {only_solved: 1.})
        else:
            c = _coef(term, X, names, xvalues)
        if term.get('outvec') is not None:
            out = out + c[:, None] * np.asarray(term['outvec'], dtype='f8')
            continue
        vec = tables[term['table']] if term.get('table') is not None else 1.
        if term.get('featvec') is not None:   # weights on P_ell(kin) BEFORE get_corr (EFT-like counter terms of a correlation-function theory)
            p2c, raw = tables['@p2c'], tables[term['table'] + '@kin'] * np.asarray(term['featvec'], dtype='f8')
            vec = np.array([p2c(r.reshape(len(p2c.ells), -1)).ravel() for r in raw])
        if term.get('colvec') is not None:
            vec = vec * term['colvec']
        P = P + c[:, None] * vec
    flat = (P @ obs['wmatrix'].T if obs.get('wmatrix') is not None else P) + out
    if only_solved is None and obs.get('offset') is not None:
        flat = flat + obs['offset']
    return flat


def _precision_blocks(spec):
    """-> list of (slice over the joint data vector, precision block) -- one entry when the precision is joint."""
    sizes = [len(obs['data']) for obs in spec['observables']]
    prec = spec['precision']
    if isinstance(prec, (list, tuple)):   # each block covers a run of consecutive observables (one, or several with a joint covariance)
        out, start = [], 0
        for pb in prec:
            pb = np.asarray(pb, dtype='f8')
            out.append((slice(start, start + pb.shape[0]), pb))
            start += pb.shape[0]
        assert start == sum(sizes), 'precision blocks do not cover the data vector'
        return out
    return [(slice(0, sum(sizes)), np.asarray(prec, dtype='f8'))]


def marginalize_precision(spec):
    """'.prec' pre-marginalisation, restated from reference desilike/likelihoods/base.py:257-312 (``_marginalize_precision``):
    once, at the fiducial point (the parameters' default values, spec['fiducial']), with J = d flattheory / d x over the '.prec'
    parameters (solved entries with mode='prec') and their Gaussian priors N(loc, scale),

        precision <- P - (J P)^T (J P J^T + diag(1 / scale^2))^-1 (J P)          (:305-308)
        flatdata  <- flatdata - sum_i loc_i J_i                                   (:309; flatdiff(x = loc) - flatdiff(x = 0))

    per likelihood (= precision block); afterwards these parameters are no longer solved: their terms stay in the theory at the
    parameters' default values.  Returns a new spec without mode='prec' entries."""
    solved = list(spec.get('solved', []))
    prec = [s for s in solved if s.get('mode') == 'prec']
    if not prec:
        return spec
    names = list(spec['params'])
    Xf = np.array([[float(spec['fiducial'][n]) for n in names]])
    x0 = {s['name']: float(s.get('value', 0.)) for s in solved}
    J = []   # [nprec][n]
    for s in prec:
        rows = []
        for obs in spec['observables']:
            if any(t.get('solved') == s['name'] for t in obs['terms']):
                rows.append(observable_flat(obs, Xf, names, x0, only_solved=s['name'])[0])
            else:
                rows.append(np.zeros(len(obs['data'])))
        J.append(np.concatenate(rows))
    J = np.array(J)
    loc = np.array([float((s.get('prior') or {}).get('loc', 0.)) for s in prec])
    scale = np.array([float((s.get('prior') or {}).get('scale', np.inf)) if (s.get('prior') or {}).get('dist') == 'norm' else np.inf for s in prec])
    prior_prec = np.where(np.isinf(scale), 0., 1. / scale**2)
    blocks = _precision_blocks(spec)
    # per LIKELIHOOD (reference :298-308): spec['prec_groups'] lists the precision blocks of each likelihood (default: one block each); the
    # update uses the likelihood's full precision, and a '.prec' parameter shared between likelihoods is an error (:285-288)
    groups = [list(g) for g in spec.get('prec_groups', [])] or [[b] for b in range(len(blocks))]
    member = [[i for i in range(len(prec)) if any(np.any(J[i, blocks[b][0]] != 0.) for b in g)] for g in groups]
    shared = [prec[i]['name'] for i in range(len(prec)) if sum(i in m for m in member) > 1]
    if shared:
        raise ValueError('cannot use .prec for parameters that are shared between likelihoods (we would need to create a joint covariance matrix!); common block is {}'.format(shared))
    new_prec, shift = [], J.T.dot(loc)
    for g, idx in zip(groups, member):
        if not idx:
            new_prec += [blocks[b][1] for b in g]; continue
        touching = [b for b in g if any(np.any(J[i, blocks[b][0]] != 0.) for i in idx)]
        if len(g) == 1 or len(touching) <= 1:
            for b in g:
                sl, P = blocks[b]
                if b not in touching: new_prec.append(P); continue
                Jb = J[idx][:, sl]
                derivp = Jb * P if P.ndim == 1 else Jb.dot(P)
                H = derivp.dot(Jb.T) + np.diag(prior_prec[idx])
                full = np.diag(P) if P.ndim == 1 else P
                new_prec.append(full - derivp.T.dot(np.linalg.solve(H, derivp)))
            continue
        sl = slice(blocks[g[0]][0].start, blocks[g[-1]][0].stop)
        import scipy.linalg
        P = scipy.linalg.block_diag(*[np.diag(blocks[b][1]) if blocks[b][1].ndim == 1 else blocks[b][1] for b in g])
        Jb = J[idx][:, sl]
        derivp = Jb.dot(P)
        H = derivp.dot(Jb.T) + np.diag(prior_prec[idx])
        new_prec.append(P - derivp.T.dot(np.linalg.solve(H, derivp)))
    out = dict(spec)
    out['solved'] = [dict(s) for s in solved if s.get('mode') != 'prec']
    out['precision'] = new_prec if isinstance(spec['precision'], (list, tuple)) else new_prec[0]
    out.pop('prec_groups', None)
    pnames = {s['name']: float(s.get('value', 0.)) for s in prec}
    observables, start = [], 0
    for obs in spec['observables']:
        o = dict(obs)
        n = len(obs['data'])
        o['data'] = np.asarray(obs['data'], dtype='f8') - shift[start:start + n]
        start += n
        terms = []
        for t in obs['terms']:
            if t.get('solved') in pnames:
                value = pnames[t['solved']]
                if value == 0.: continue
                t = dict(t, solved=None, scale=float(t.get('scale', 1.)) * value)
            terms.append(t)
        o['terms'] = terms
        observables.append(o)
    out['observables'] = observables
    return out


def evaluate(spec, X, return_theory=False, return_hessian=False):
    """X [N, d] -> dict(logposterior, loglikelihood, logprior, x_solved [N, ns], flattheory [N, n][, hessian [N, ns, ns]])."""
    spec = marginalize_precision(spec)
    X = np.atleast_2d(np.asarray(X, dtype='f8'))
    names = list(spec['params'])
    if spec.get('derived'):
        # derived-expression parameters (reference parameter.py:795-808 Parameter.eval -> utils.evaluate(expression, locals); applied per point at
        # base.py:537): evaluated with Python's eval on the batch columns, like the reference; the prior stays over the sampled parameters
        cols = {n: X[:, i] for i, n in enumerate(names)}
        for dp in spec['derived']:
            import re
            deps = re.findall(r'\{([^{}]+)\}', dp['expr'])
            expr = dp['expr']
            for j, n in enumerate(deps): expr = expr.replace('{' + n + '}', '__p{:d}__'.format(j))
            cols[dp['name']] = np.broadcast_to(eval(expr, {'np': np, 'jnp': np, 'sqrt': np.sqrt, 'exp': np.exp, 'log': np.log, 'pi': np.pi, '__builtins__': {}},
                                                    {'__p{:d}__'.format(j): cols[n] for j, n in enumerate(deps)}), (X.shape[0],)).astype('f8')
        sampled = list(names)
        names = sampled + [dp['name'] for dp in spec['derived']]
        X = np.column_stack([cols[n] for n in names])
        spec = dict(spec, params=names, derived=[], priors=dict(spec.get('priors', {}), **{dp['name']: None for dp in spec['derived']}))
    solved = spec.get('solved', [])
    x0 = {s['name']: float(s.get('value', 0.)) for s in solved}
    N = X.shape[0]
    flats, jacs = [], []
    for obs in spec['observables']:
        tables = theory_tables(obs['theory'], X, names)
        flat = observable_flat(obs, X, names, x0, tables=tables)
        dscale = 1.
        if obs.get('transform') == 'cubic':
            # the reference obtains the Jacobian of flatdiff by jacfwd THROUGH the transform (likelihoods/base.py:152-158), at the solved
            # parameters' defaults: rows u'(t) dt/dx with u'(t) = (t / d)^(-2/3) for u = (3 (t / d)^(1/3) - 2) d (power_spectrum.py:402-404)
            dscale = np.cbrt(flat / obs['data'])**-2
            flat = core.cubic_transform(flat, obs['data'])
        flats.append(flat)
        if solved:
            snames = {t.get('solved') for t in obs['terms']}
            jacs.append(np.stack([dscale * observable_flat(obs, X, names, x0, tables=tables, only_solved=s['name']) if s['name'] in snames
                                  else np.zeros_like(flat) for s in solved], axis=1))  # [N, ns, nobs]
    flattheory = np.concatenate(flats, axis=-1)
    flatdata = np.concatenate([obs['data'] for obs in spec['observables']])
    flatdiff = flattheory - flatdata
    blocks = _precision_blocks(spec)
    logprior = core.logprior(X, names, spec.get('priors', {}))
    out = {}
    if not solved:
        loglike = np.zeros(N)
        for sl, P in blocks:
            loglike += -0.5 * core.chi2(flatdiff[:, sl], P)
        out['x_solved'] = np.zeros((N, 0))
    else:
        J = np.concatenate(jacs, axis=-1)  # [N, ns, n]
        loglike, xs = np.empty(N), np.empty((N, len(solved)))
        lp_solved = np.empty(N)
        priors = [s.get('prior') for s in solved]
        marg = [bool(s.get('marg', True)) for s in solved]
        xx0 = [x0[s['name']] for s in solved]
        hess = np.empty((N, len(solved), len(solved)))
        for ip in range(N):
            res = core.solve_marginalised([flatdiff[ip, sl] for sl, P in blocks], [J[ip][:, sl] for sl, P in blocks], [P for sl, P in blocks],
                                          xx0, priors, marg, with_hessian=True)
            loglike[ip], xs[ip], lp_solved[ip], hess[ip] = res['loglikelihood'], res['x'], res['logprior_solved'], res['hessian']
        if return_hessian: out['hessian'] = hess
        logprior = logprior + lp_solved
        out['x_solved'] = xs
    out.update(loglikelihood=loglike, logprior=logprior, logposterior=loglike + logprior)
    if return_theory:
        out['flattheory'] = flattheory
    return out
