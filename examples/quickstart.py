"""Quick start (needs a B200 and the built library: ``python -m desilike_b200.build``).

The same objects and calls as the reference (desilike docs/user/getting_started.rst; nb/fs_shapefit_examples.ipynb), with the batch evaluated
on the GPU.  Every call below is exercised by tests/test_gpu_api.py and tests/test_sampler.py.
"""
import numpy as np

import desilike_b200 as dl


def pk_fid(k):   # any tabulated / callable fiducial linear power spectrum (cosmoprimo provides it in a real analysis)
    return 2e4 * (k / 0.02) / (1. + (k / 0.02)**2.6) * (1. + 0.05 * np.sin(k * 105.) * np.exp(-(k * 8.)**2))


# theory -> observable -> likelihood, as in desilike
template = dl.ShapeFitPowerSpectrumTemplate(pk_dd_fid=pk_fid, pknow_dd_fid=pk_fid, f_fid=0.8)
theory = dl.EFTLikeKaiserTracerPowerSpectrumMultipoles(template=template)
for param in theory.params.select(basename=['ct*', 'sn*']):   # counter / stochastic terms: analytic marginalisation
    param.update(derived='.marg')
k = np.arange(0.02, 0.2, 0.005)
observable = dl.TracerPowerSpectrumMultipolesObservable(data={'b1': 2.}, k=[k] * 3, ells=(0, 2, 4), theory=theory, shotnoise=1e4,
                                                        covariance=np.diag(np.full(3 * k.size, 1e5)))
likelihood = dl.ObservablesGaussianLikelihood([observable])

print('single point:', likelihood(b1=1.9, qpar=1.01))          # calculator-style call (reference base.py:124-128)

# a batch of points: one CUDA launch sequence for all of them (reference base.py:232, backend None / 'jax' / 'mpi')
rng = np.random.default_rng(0)
params = {'b1': rng.uniform(1.5, 2.5, 65536), 'qpar': rng.uniform(0.95, 1.05, 65536), 'qper': rng.uniform(0.95, 1.05, 65536)}
logpost, derived = dl.vmap(likelihood, backend='cuda', return_derived=True)(params)
print('batch:', logpost.shape, 'best:', logpost.max(), 'marginalised sn0 at the best point:', derived[likelihood.plan.solved_names[3]][np.argmax(logpost)])

# device-resident ensemble sampler (emcee's stretch move; walkers, proposals and log-posteriors never leave the GPU)
plan = likelihood.plan
nwalkers = 4096
start = np.column_stack([rng.normal(p.value, 0.01, nwalkers) for p in likelihood.varied_params])
sampler = dl.EnsembleSampler(plan, nwalkers, seed=1)
sampler.set_start(start)
sampler.run(200, store=False)
chain, logp = sampler.run(100, thin_by=10)
print('chain:', tuple(chain.shape), 'acceptance fraction:', sampler.acceptance_fraction)
