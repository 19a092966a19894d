"""Shared construction of a velocileptors-style (pktable) Taylor-emulated likelihood for the CPU and GPU tests."""
import itertools

import numpy as np


def taylor_powers(d, order):
    return np.array([p for total in range(order + 1) for p in itertools.product(range(order + 1), repeat=d) if sum(p) == total], dtype='i4')


def make_lpt_likelihood(marg=True, seed=11, order=2, nk=24, tracer=None, engine='taylor', physical=False):
    """-> (likelihood, emulator-state pieces) with a random Taylor emulator of pktable [3, nk_pt, 19] in (qpar, qper, dm, df)."""
    import desilike_b200 as dl
    from oracle import synthetic
    rng = np.random.default_rng(seed)
    ells = (0, 2, 4)
    kpt = np.concatenate([[0.0005], np.geomspace(0.0015, 0.025, 20, endpoint=True), np.arange(0.03, 0.5 + 0.0075, 0.005)])   # reference full_shape.py:1283
    k = np.linspace(0.02, 0.2, nk)
    params = ['qpar', 'qper', 'dm', 'df']
    center = np.array([1., 1., 0., 1.])
    powers = taylor_powers(len(params), order)
    base = synthetic.pk_fid(kpt)
    amp = np.array([1., 1., 1., 0.3, 0.3, 0.1, 0.3, 0.3, 0.1, 0.1, 0.2, 0.2, 0.05, 0.05, 0.05, 0.05, 1e-4, 1e-4, 1e-4])
    kfac = np.ones((len(kpt), 19)); kfac[:, 12:16] = (kpt**2)[:, None] * 10.; kfac[:, 16:] = np.stack([np.ones_like(kpt), kpt**2, kpt**4], axis=-1) / base[:, None]
    deriv = rng.standard_normal((len(powers), len(ells), len(kpt), 19)) * (0.2**powers.sum(axis=-1))[:, None, None, None] * base[None, None, :, None] * (amp * kfac)[None, None]
    deriv[0] = np.abs(deriv[0]) + base[None, :, None] * (amp * kfac)[None] * np.array([1., 0.5, 0.2])[:, None, None]
    all_params = [dl.Parameter('qpar', value=1., prior=dict(limits=[0.8, 1.2])), dl.Parameter('qper', value=1., prior=dict(limits=[0.8, 1.2])),
                  dl.Parameter('dm', value=0., prior=dict(limits=[-0.5, 0.5])), dl.Parameter('df', value=1., prior=dict(limits=[0.5, 1.5]))]
    mlp = None
    if engine == 'mlp':   # a small random network whose output is the flat pktable [3, nk_pt, 19] (affine y-scaling onto the table amplitudes)
        ny = len(ells) * len(kpt) * 19
        widths = [len(params), 12, 10, ny]
        layers = [(rng.standard_normal((a, b)) * np.sqrt(2. / a), 0.1 * rng.standard_normal(b)) for a, b in zip(widths[:-1], widths[1:])]
        scale = np.abs(deriv[0]).ravel() + 1e-3
        mlp = dict(xlimits=np.array([[0.8, 1.2], [0.8, 1.2], [-0.5, 0.5], [0.5, 1.5]]), layers=layers, activation='silu',
                   ylimits=np.stack([deriv[0].ravel() - 0.05 * scale, deriv[0].ravel() + 0.05 * scale], axis=-1), pca=None)
        emu = dl.Emulator.from_mlp(params, mlp, {'pktable': slice(0, ny)}, {'pktable': (len(ells), len(kpt), 19)}, fixed=dict(k=kpt, ells=ells, z=1.), all_params=all_params,
                                   defaults=dict(zip(params, center)), calculator_class=('desilike.theories.galaxy_clustering.full_shape', 'LPTVelocileptorsPowerSpectrumMultipoles'))
    else:
        tables = {'pktable': deriv}
        scalars = None
        if physical:   # sigma8(theta), fsigma8(theta) as scalar Taylor engines of the same emulator (what the pt calculator exposes, reference full_shape.py:1300-1301)
            s8 = rng.standard_normal(len(powers)) * 0.05 * 0.3**powers.sum(axis=-1); s8[0] = 0.8
            fs8 = rng.standard_normal(len(powers)) * 0.03 * 0.3**powers.sum(axis=-1); fs8[0] = 0.45
            tables.update(sigma8=s8, fsigma8=fs8)
            scalars = dict(sigma8=s8, fsigma8=fs8)
        emu = dl.Emulator.from_taylor(params, center, powers, tables, fixed=dict(k=kpt, ells=ells, z=1.), all_params=all_params,
                                      calculator_class=('desilike.theories.galaxy_clustering.full_shape', 'LPTVelocileptorsPowerSpectrumMultipoles'))
    theory = dl.LPTVelocileptorsTracerPowerSpectrumMultipoles(pt=emu.to_calculator(), k=k, ells=ells, shotnoise=0.8e4 if physical else 1e4, tracer=tracer,
                                                              prior_basis='physical' if physical else 'standard', tracer_type='LRG' if physical else None)
    if marg:
        for name in ['alpha0', 'alpha2', 'alpha4', 'sn0', 'sn2', 'sn4']:
            theory.params[theory._ns(name + ('p' if physical else ''))].update(derived='.marg')
    n = len(ells) * nk
    fid = dict(b1=1.8, b2=0.4, bs=-0.3)
    # data = the model at a fiducial point (linear parameters at 0) + noise
    m = np.zeros(19); m[[0, 1, 2, 3, 4, 5, 6, 7, 8, 9]] = [1., fid['b1'], fid['b1']**2, fid['b2'], fid['b1'] * fid['b2'], fid['b2']**2, fid['bs'], fid['b1'] * fid['bs'], fid['b2'] * fid['bs'], fid['bs']**2]
    from scipy import interpolate
    data = interpolate.interp1d(kpt, (deriv[0] * m).sum(axis=-1), kind='cubic', axis=-1)(k).ravel()
    sig = 0.05 * np.abs(data) + 20.
    A = rng.standard_normal((n, n // 3)); C = A.dot(A.T); C /= np.sqrt(np.outer(np.diag(C), np.diag(C)))
    cov = (0.85 * np.eye(n) + 0.15 * C) * np.outer(sig, sig)
    data = data + 0.3 * sig * rng.standard_normal(n)
    obs = dl.TracerPowerSpectrumMultipolesObservable(data=data.reshape(len(ells), nk), k=[k] * len(ells), ells=ells, theory=theory, name='lpt')
    like = dl.ObservablesGaussianLikelihood(observables=[obs], covariance=cov)
    info = dict(kpt=kpt, k=k, ells=ells, params=params, center=center, powers=powers, deriv=deriv, cov=cov, data=data, mlp=mlp)
    if physical: info.update(scalars=scalars, physical=dict(shotnoise=0.8e4, fsat=theory.fsat, sigv=theory.sigv))
    return like, info


def direct_theory(info, X, names, nd=1e-4):
    """Un-collapsed restatement through the oracle: Taylor predict of pktable -> bias combine -> cubic re-interpolation."""
    from oracle import core
    out = []
    for x in X:
        v = dict(zip(names, x))
        if info.get('mlp') is not None:
            pktable = core.mlp_predict(np.array([[v[p] for p in info['params']]]), info['mlp'])[0].reshape(len(info['ells']), len(info['kpt']), 19)
        else:
            diffs = np.array([v[p] for p in info['params']]) - info['center']
            mono = np.prod(diffs**info['powers'], axis=-1)
            pktable = np.tensordot(mono, info['deriv'], axes=(0, 0))
        pars = [v.get(n, 0.) for n in core.LPT_BIAS_PARAMS]
        if info.get('physical') is not None:   # physical prior basis (reference full_shape.py:1299-1307): sigma8 / fsigma8 from their Taylor engines
            sigma8, fsigma8 = mono.dot(info['scalars']['sigma8']), mono.dot(info['scalars']['fsigma8'])
            p = {n: v.get(n, 0.) for n in ['b1p', 'b2p', 'bsp', 'b3p', 'alpha0p', 'alpha2p', 'alpha4p', 'sn0p', 'sn2p', 'sn4p']}
            pars = core.lpt_physical_pars(p, sigma8, fsigma8, **info['physical'])
        out.append(core.lpt_tracer_power(pktable, info['kpt'], info['k'], pars, nd=nd).ravel())
    return np.array(out)
