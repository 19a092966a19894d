#!/usr/bin/env python
"""Benchmark of the batched full-shape likelihood path (BASELINE.json metric: logL evaluations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c5|c5solved|c1|c3|c3fft|c4wide|c5step] [--batch B]
                    [--impl ours|reference] [--extras auto|none|c2,c5,...] [--sustained S]

A *step* is one pass of the hot path over one batch of ``--batch`` synthetic parameter points PER GPU.  Default workload = the
5-tracer emulated configuration BASELINE.json's target is stated on (configs[3], SURVEY.md section 8d C4: 5 tracers x [MLP pt emulator ->
EFT-like combine -> window] -> joint chi2 with analytic marginalisation of 35 counter / stochastic terms), 131 072 points per GPU -- at
8 GPUs exactly the "1M params on 8xB200" of configs[3]; the same workload at every N (weak scaling).  Prints ONE JSON line (rank 0).

  value      device-resident throughput: inputs already in HBM, K steps bracketed by barrier + synchronize, CUDA events, max over ranks.
             Inputs rotate through a pool of distinct batches larger than L2 (no L2 flush needed).  With N > 1 GPUs every step also
             delivers logposterior and the marginalised nuisance values of ALL ranks' points to EVERY rank: the finish kernel stores them
             straight into the peers' gather windows over NVLink (csrc/gather.cu; NCCL all-gather when CUDA IPC is unavailable), and the
             gathered window is checked against a single-GPU evaluation after the timed region (`gather_check`).
  sustained  the same steps back to back for >= --sustained seconds (default 2), clocks sampled during it.
  e2e        same metric through the public host API (pinned HOST buffers): H2D of the parameters and D2H of the results inside the
             timed region.
  roofline   dominant kernel: algorithmic FLOPs of the folded chain per launch / CUDA-event time of those launches, against the FP64
             DMMA peak measured in this run (MEASURED_PEAKS.json has no FP64 entry; a cuBLAS DGEMM 8192^3 figure is reported beside it).
  cpu_baseline  the NumPy oracle port (threaded BLAS) on a bounded sample of the same workload on this box's host cores; `folded`: the
             same plan-time-folded algebra the GPU executes, in vectorised NumPy (so that the GPU/CPU ratio is not read as pure hardware).
  extra      compact lines (value, e2e, roofline, cpu_baseline) of the other SURVEY.md section 8d workloads, measured in the same run.

Synthetic inputs: the workload definitions (spec constants, fiducial data vector, covariance) of SURVEY.md section 8d live product-side in
desilike_b200/workloads.py (the tests reach the same generators through oracle/configs.py).  The default arm builds them without importing
oracle/ (the synthetic data vector = theory at a fiducial point comes from the package's own CUDA path); every timed region of the default
arm executes desilike_b200 (the C-ABI library) only; oracle/pipeline.py runs in the cpu_baseline leg and in ``--impl reference`` (whose
workloads are built through oracle/configs.py: no GPU needed).

``--impl reference``: the reference is pure Python and /root/reference does not travel to the GPU box, so the reference
arm times the oracle port (oracle/pipeline.py -- the one other place bench.py may execute oracle/) with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

# stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, ...) write to file descriptor 1 directly, so the real
# stdout is kept aside for the result line and fd 1 is pointed at stderr for everything else
_RESULT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)
if 'reference' in sys.argv:   # the CPU arm uses every host core (torchrun exports OMP_NUM_THREADS=1); must precede the NumPy import
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count())

import numpy as np


def emit(line):
    _RESULT.write(json.dumps(line) + '\n')
    _RESULT.flush()


def log(*args):
    print('[bench]', *args, file=sys.stderr, flush=True)


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'logL evals/sec (batched, emulated FS+window+chi2)'
# algorithmic FLOP per evaluation of the UN-collapsed chain
UNCOLLAPSED_FLOPS = {'c1': 0.1e6, 'c2': 0.7e6, 'c3': 1.3e6, 'c3fft': 1.3e6, 'c4': 20e6, 'c4wide': 47e6, 'c4wide_i8': 47e6, 'c4dense': 300e6, 'c5': 8e6, 'c5solved': 19e6}
UNIT = 'evals/s'
DEFAULT_WORKLOAD, DEFAULT_BATCH = 'c4', 131072
EXTRA_BATCH = {'c4wide': 16384, 'c4wide_i8': 16384, 'c4dense': 16384}
L2_BYTES = 126 * 2**20


_IMPL = 'ours'


def workload_module():
    """Synthetic-workload generators.  Default arm: the product-side module (desilike_b200/workloads.py; the synthetic data vectors -- theory at a
    fiducial point -- come from the package's own CUDA path).  Reference arm: the same generators through oracle/configs.py, which installs the oracle
    for that step (no GPU needed)."""
    if _IMPL == 'reference':
        from oracle import configs
        return configs
    from desilike_b200 import workloads
    return workloads


def build_spec(workload):
    configs = workload_module()
    if workload == 'c1':
        return configs.config_c1(), 'C1: KaiserTracerPowerSpectrumMultipoles (ShapeFit template, ell=0,2,4, k=0.02-0.2, 36 bins/ell) -> chi2 (n=108, d=6), direct theory (no emulator)'
    if workload == 'c3':
        return configs.config_c3(), 'C3: damped-BAO P(k) -> FFTLog (2048->4096) -> xi_ell(s) + solved broadband -> chi2 (n=66, d=7, ns=4); P -> xi operator folded into the term matrices at plan time'
    if workload == 'c3fft':
        spec = configs.config_c3()
        spec['observables'][0]['theory']['fftlog'] = 'kernel'
        return spec, 'C3 with the explicit FFTLog kernel: damped-BAO P(k) -> fftlog_kernel (2048->4096) -> xi_ell(s) + solved broadband -> chi2 (n=66, d=7, ns=4)'
    if workload == 'c2':
        return configs.config_c2(coefficients='random'), 'C2: TaylorEmulator(order 4, T=70) pt tables 3x[3,350] -> Kaiser combine -> window [108,1050] -> chi2 (n=108, d=6)'
    if workload == 'c5':
        return configs.config_c5(coefficients='random'), 'C5: 5 tracers x [Taylor(order 4) -> combine -> window [402,1050]] -> block-diagonal chi2 (n=2010, d=14)'
    if workload == 'c5solved':
        return configs.config_c5(coefficients='random', solved=True), 'C5 + analytic marginalisation of 35 ct/sn terms (n=2010, d=9, ns=35)'
    if workload == 'c4':
        return configs.config_c4(), 'C4: 5 tracers x [MLP 4-64-64-64 pt emulator -> EFT-like combine -> window [402,1050]] -> joint chi2 + analytic marginalisation (n=2010, d=25, ns=35)'
    if workload == 'c4dense':
        return configs.config_c4(dense=True), 'C4 dense stress variant: 5 tracers x [MLP 4-64-64-64 -> EFT-like combine -> window [402,1050]] -> chi2 with a DENSE joint 2010 x 2010 precision + marginalisation of 35 terms (15 theta-dependent rows in one block: split over passes of 8 slots)'
    if workload == 'c4wide_i8':
        spec, label = build_spec('c4wide')
        return spec, label + '; hidden layers as FP64-grade GEMMs on the tcgen05 int8 tensor cores (6 exact slices, opt-in mlp_slices)'
    if workload == 'c4wide':
        return configs.config_c4(hidden=(512, 512, 512, 512)), 'C4 stress variant: 5 tracers x [MLP 4-512-512-512-512 pt emulator -> EFT-like combine -> window [402,1050]] -> chi2 + marginalisation (ns=35)'
    raise ValueError(workload)


def pool_size(N, d):
    return max(2, int(np.ceil(1.25 * L2_BYTES / (N * d * 8))))


def config_dict(label, N, world, d):
    """Identical for the GPU arm and the reference arm (the driver compares them)."""
    npool = pool_size(N, d)
    return {'workload': label, 'batch_per_gpu': N,
            'parallelism': 'points sharded over {:d} GPU(s); logposterior + marginalised nuisance values of all points delivered to every rank'.format(world) if world > 1 else 'single GPU',
            'l2_policy': 'inputs rotate through a pool of {:d} distinct batches ({:.0f} MB > 126 MB L2)'.format(npool, npool * N * d * 8 / 2**20)}


def host_batch(spec, N, rank, idx):
    """Batch ``idx`` of rank ``rank``'s input pool (deterministic: any rank can rebuild any other rank's inputs for the gather check)."""
    configs = workload_module()
    X = configs.sample_points(spec, N, seed=1000 + 97 * rank + idx % 8)
    if idx >= 8: X = X * (1. + 1e-6 * idx)   # cheap distinct batches: a jitter within the prior box keeps the workload identical
    return X


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None: return
        nv = self.nvml
        names = {getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown', getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown', getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try: mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception: mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit: self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        if self.nvml is None or not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ---- CPU legs ---------------------------------------------------------------------------------------------------------------------------
def cpu_port(spec, X, seconds=10., max_points=None):
    """Oracle port (vectorised NumPy, threaded BLAS) on a bounded sample; returns (evals/s, points evaluated)."""
    from oracle import pipeline
    n = min(len(X), max_points or len(X))
    pipeline.evaluate(spec, X[:min(n, 256)])  # warm up (imports, BLAS threads)
    done, t0 = 0, time.perf_counter()
    while True:
        pipeline.evaluate(spec, X[:n])
        done += n
        dt = time.perf_counter() - t0
        if dt >= seconds: break
    return done / dt, done


def cpu_folded(low, X):
    """The plan-time-FOLDED chain in vectorised NumPy: the same per-point algebra the GPU kernels execute on the lowered plan
    (features -> bias monomials -> term contraction -> |r|^2 / normal equations -> prior), i.e. what a CPU program gets from the same
    folding.  Only Taylor and MLP generators (the emulated workloads); returns logposterior [N]."""
    N = X.shape[0]
    F = []
    for gen in low.gens:
        cols = [X[:, int(gen.param_idx[j])] if int(gen.param_idx[j]) >= 0 else np.full(N, gen.param_default[j]) for j in range(len(gen.params))]
        Xe = np.stack(cols, axis=-1)
        if gen.kind_name == 'taylor':
            T = gen.nfeat - 2
            powers = np.asarray(gen.idata).reshape(T, len(gen.params))
            diffs = Xe - np.asarray(gen.ddata)
            pw = [np.stack([diffs[:, j]**e for e in range(int(powers[:, j].max()) + 1)], axis=-1) for j in range(diffs.shape[1])]
            mono = np.ones((N, T))
            for j in range(diffs.shape[1]): mono *= pw[j][:, powers[:, j]]
            F.append(np.concatenate([np.ones((N, 1)), np.zeros((N, 1)), mono], axis=-1))
        elif gen.kind_name == 'mlp':
            idata, dd = np.asarray(gen.idata), np.asarray(gen.ddata)
            nh, act, d = int(idata[0]), int(idata[1]), int(idata[2])
            v = Xe * dd[:d] + dd[d:2 * d]
            o = 2 * d
            widths = [d] + [int(w) for w in idata[3:3 + nh]]
            for l in range(nh):
                W = dd[o:o + widths[l] * widths[l + 1]].reshape(widths[l], widths[l + 1]); o += W.size
                b = dd[o:o + widths[l + 1]]; o += b.size
                v = v @ W + b
                v = v / (1. + np.exp(-v)) if act == 1 else (np.maximum(v, 0.) if act == 2 else (np.tanh(v) if act == 3 else v))
            F.append(np.concatenate([np.ones((N, 1)), np.zeros((N, 1)), v], axis=-1))
        else:
            raise ValueError('cpu_folded: emulated workloads only')
    C = np.ones((N, max(len(low.coef_keys), 1)))
    for key, j in low.coef_keys.items():
        for idx, power in key: C[:, j] *= X[:, idx]**power
    ns = low.ns
    skind, sloc, sscale, slo, shi = low.xprior
    xprec = np.where(skind == 2, 1. / sscale**2, 0.)
    M = np.zeros((N, ns, ns)) + np.diag(xprec)
    gL, chi2 = np.zeros((N, ns)), np.zeros(N)
    for blk in low.blocks:
        slots = np.zeros((blk['nslots'], N, blk['ncols_white']))
        for (slot, coef, gen, foff, K), mw in zip(blk['terms'], blk['white']):
            slots[slot] += (C[:, coef][:, None] if coef >= 0 else 1.) * (F[gen][:, foff:foff + K] @ mw)
        r = slots[0] - blk['r0_white']
        chi2 += np.einsum('ij,ij->i', r, r)
        rows = [(int(si), slots[1 + a]) for a, si in enumerate(blk['slot_solved'])]
        crow = [(int(si), blk['gconst_white'][i]) for i, si in enumerate(blk['const_solved'])]
        for si, Ji in rows:
            gL[:, si] -= np.einsum('ij,ij->i', Ji, r)
            for sj, Jj in rows: M[:, si, sj] += np.einsum('ij,ij->i', Ji, Jj)
            for sj, gj in crow:
                v = Ji @ gj
                M[:, si, sj] += v; M[:, sj, si] += v
        for si, gi in crow:
            gL[:, si] -= r @ gi
            for sj, gj in crow: M[:, si, sj] += gi.dot(gj)
    kind, lo, hi, loc, scale = low.prior
    lp = np.zeros(N)
    for j in range(low.d):
        isin = (lo[j] <= X[:, j]) & (X[:, j] <= hi[j])
        lp += np.where(isin, -0.5 * ((X[:, j] - loc[j]) / scale[j])**2 if kind[j] == 2 else 0., -np.inf)
    ll = -0.5 * chi2
    if ns:
        g = -(low.x0 - sloc) * xprec + gL
        Lc = np.linalg.cholesky(M)
        dx = np.linalg.solve(M, g[..., None])[..., 0]
        ll += 0.5 * (-np.einsum('ij,ij->i', g, dx) + np.einsum('j,ij->i', xprec, dx**2)) + np.einsum('ij,ij->i', gL, dx)
        ll -= np.sum(np.log(np.diagonal(Lc, axis1=1, axis2=2)[:, :low.nmarg]), axis=-1)
        x = low.x0 + dx
        lp += np.sum(np.where(skind == 2, -0.5 * (x - sloc)**2 * xprec, 0.), axis=-1)
    return ll + lp


def cpu_folded_rate(low, X, seconds):
    n = min(len(X), 4096)
    cpu_folded(low, X[:256])
    done, t0 = 0, time.perf_counter()
    while True:
        cpu_folded(low, X[:n]); done += n
        dt = time.perf_counter() - t0
        if dt >= seconds: break
    return done / dt, done


def recorded_reference_loop():
    """The reference's own serial `vmap(likelihood, backend=None)` loop cannot run on the GPU box (/root/reference does not travel): its
    timings in the build container are shipped as a recorded constant (scripts/time_reference_loop.py -> profiles/reference_loop.json)."""
    path = os.path.join(ROOT, 'profiles', 'reference_loop.json')
    try: return json.load(open(path))
    except Exception: return None


def run_reference(args, spec, label, X):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import pipeline
    pipeline.evaluate(spec, X[:256])
    t0 = time.perf_counter()
    n0 = min(1024, len(X))
    pipeline.evaluate(spec, X[:n0])
    speed = n0 / (time.perf_counter() - t0)
    # bounded sample per step: the whole --steps K run stays within ~1-2 minutes of CPU time
    sample = int(min(len(X), max(256, speed * 60. / max(args.steps + min(args.warmup, 2), 1))))
    for _ in range(min(args.warmup, 2)): pipeline.evaluate(spec, X[:sample])
    t0 = time.perf_counter()
    for _ in range(args.steps): pipeline.evaluate(spec, X[:sample])
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cores = os.cpu_count()
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(label, args.batch, args.gpus, len(spec['params'])),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '{:d} points/step x {:d} steps (each step is a bounded sample of the GPU batch), oracle/pipeline.py (vectorised NumPy, threaded BLAS)'.format(sample, args.steps),
                             'reference_loop_recorded': recorded_reference_loop()},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    emit(line)


# ---- GPU legs ---------------------------------------------------------------------------------------------------------------------------
class Context(object):

    def __init__(self):
        import torch
        self.torch = torch
        self.rank, self.world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
        self.local_rank = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        self.dist = self.comm = None
        if self.world > 1:
            import torch.distributed as dist
            from desilike_b200.distributed import Communicator
            dist.init_process_group('nccl', device_id=self.dev)
            self.dist = dist
            self.comm = Communicator(peer_gather=not os.environ.get('DLB_BENCH_NCCL'))

    def max_over_ranks(self, value):
        if self.world == 1: return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.world > 1: self.dist.barrier()
        self.torch.cuda.synchronize()


def fp64_peaks(ctx):
    """FP64 roofline denominators measured in this run: own DMMA / DFMA register loops and an independent cuBLAS DGEMM 8192^3."""
    import desilike_b200
    torch = ctx.torch
    peaks = {'dmma': desilike_b200.measure_fp64_peak('dmma'), 'dfma': desilike_b200.measure_fp64_peak('dfma')}
    try:
        n = 8192
        a = torch.randn((n, n), dtype=torch.float64, device=ctx.dev); b = torch.randn((n, n), dtype=torch.float64, device=ctx.dev)
        torch.matmul(a, b); torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peaks['dgemm_cublas_8192'] = 2. * n**3 / (best * 1e-3) / 1e12
        del a, b
        torch.cuda.empty_cache()
    except Exception as exc:   # noqa: BLE001
        peaks['dgemm_cublas_8192'] = None
        log('cuBLAS DGEMM peak unavailable:', exc)
    return peaks


def roofline_for(workload, plan, phase, N, value_per_gpu, peaks):
    flops = plan.lowered.flops_per_point()
    step_ms = phase['features'] + phase['contract'] + phase['solve']
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'traffic_{}.json'.format(workload))
    if os.path.exists(prof):
        try: traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        except Exception: traffic = None
    src = 'measured in this run; MEASURED_PEAKS.json has no FP64 entry (cuBLAS DGEMM 8192^3 in this run: {} TFLOP/s)'.format(
        None if peaks.get('dgemm_cublas_8192') is None else round(peaks['dgemm_cublas_8192'], 2))
    if workload == 'c3':
        # dominant kernel: damped-BAO P_ell(kin) (one warp per point; FP64 FMA pipe).  Algorithmic work (SURVEY.md section 8a row a12,
        # 8d "BAO P(k)"): nk * nmu nodes x (3 cubic interpolations ~ 20 FLOP each + damping exp ~ 30 + AP / Kaiser / FoG factors ~ 40)
        # + the Legendre sums 2 nell nk nmu, with nk = 300, nmu = 10, nell = 2
        nk_, nmu_, nell_ = 300, 10, 2
        kflops, kname, kms, peak = nk_ * nmu_ * (3 * 20 + 30 + 40) + 2 * nell_ * nk_ * nmu_, 'dlb::bao_kernel (FP64 FMA pipe; AP + wiggle spline + Gauss-Legendre quadrature)', phase['features'], peaks['dfma']
    elif workload == 'c3fft':
        # dominant kernel: FFTLog; SURVEY.md section 8d: n_ell (2 * 2.5 Np log2 Np + 6 Np) FLOP per evaluation with Np = 4096
        nell, Np = 2, 4096
        kflops, kname, kms, peak = nell * (2 * 2.5 * Np * np.log2(Np) + 6 * Np), 'dlb::fftlog_kernel (FP64 FMA pipe, radix 16x16x8 passes in shared memory)', phase['fftlog'], peaks['dfma']
    elif workload == 'c1':
        # dominant kernel: Kaiser + ShapeFit + AP.  SURVEY.md section 8d: Nkin * nmu * 60 + 500 * (exp, tanh) (special functions at ~30 FLOP each)
        kflops, kname, kms, peak = 36 * 8 * 60 + 500 * 2 * 30, 'dlb::kaiser_kernel (FP64 FMA pipe; ShapeFit + spline + Gauss-Legendre quadrature)', phase['features'], peaks['dfma']
    elif workload.endswith('_i8') and plan.kernels.get('int8_contraction'):
        return roofline_int8(workload, plan, phase, N, flops, traffic, peaks)
    else:
        kflops, kname, kms, peak = flops['contract'] + flops['gram'], 'dlb::contract_kernel (FP64 DMMA.8x8x4, persistent, one launch per likelihood{})'.format(
            '; quadratic-form plan: features x [K, T (T + 1) / 2 K + 8 T] instead of T matrices [K, n]' if plan.kernels.get('quadratic_form') else ''), phase['contract'], peaks['dmma']
    achieved = kflops * N / (kms * 1e-3) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic, 'kernel': kname,
                'launches_per_step': 1, 'ms_per_step_in_kernel': kms, 'flops_per_eval_kernel': kflops, 'flops_per_eval_folded': flops,
                'peak_source': ('dlb_measure_fp64_peak(DMMA.8x8x4 register loop) ' if peak == peaks['dmma'] else 'dlb_measure_fp64_peak(DFMA register loop) ') + src,
                'peak_dgemm_cublas_8192': peaks.get('dgemm_cublas_8192'), 'share_of_step': kms / step_ms if step_ms else None, 'phase_ms': phase}
    if workload not in ('c1', 'c3', 'c3fft'):   # what the kernel multiplies (padding inside tiles included) next to the algorithmic count above
        roofline['flops_per_eval_executed'] = plan.executed_flops()
        roofline['achieved_executed'] = plan.executed_flops() * N / (phase['contract'] * 1e-3) / 1e12
        roofline['frac_executed'] = roofline['achieved_executed'] / peaks['dmma']
    uncollapsed = UNCOLLAPSED_FLOPS.get(workload)
    if uncollapsed:   # the reference's un-collapsed chain: an ALGEBRAIC factor of the plan-time folding, not a roofline fraction
        roofline['flops_per_eval_uncollapsed'] = uncollapsed
        roofline['algebraic_folding_factor'] = uncollapsed / max(flops['total'], 1)
    return roofline


def roofline_int8(workload, plan, phase, N, flops, traffic, peaks):
    """Wide-MLP plans on the tcgen05 int8 path (csrc/ozaki.cu): the dominant kernel is ozaki_gemm_kernel (hidden layers + one GEMM per likelihood block).
    Algorithmic count = int8 multiply-adds the exact slicing scheme needs: ns (ns + 1) / 2 slice products of every [128 x 64 x 32] tile, i.e.
    2 * ns (ns + 1) / 2 * sum over the GEMMs of Kpad * Ncols (padded to the 32-deep k blocks / 64-wide column tiles) per point.  Denominator: int8
    dense = 2 x the bf16 dense throughput of MEASURED_PEAKS.json (the tensor pipe runs int8 at twice the bf16 rate; nominal 4500 TOP/s).  Time:
    CUDA events around the features + contraction phases of the step, i.e. INCLUDING the slicing, first-layer and Gram-reduction kernels."""
    ns = plan.lowered.mlp_slices
    # GEMM shapes from the lowered generators (idata of an MLP generator = [nhidden, act, d, H_1 ..]) / blocks
    gemm = []
    for g in plan.lowered.gens:
        if g.kind_name == 'mlp':
            idata = [int(v) for v in g.idata]
            widths = [idata[2]] + idata[3:3 + idata[0]]
            for a, b in zip(widths[:-1], widths[1:]):
                if a >= 64: gemm.append((a, b))
    for b in plan.lowered.blocks:
        K = b['terms'][0][4]
        ncolx = (b['ncols_white'] or b['ncols']) + (0 if b['gconst_white'] is None else len(b['gconst_white']))
        ncolx += ncolx & 1
        gemm.append((K, len(b['terms']) * ncolx))
    pad = lambda n, m: (n + m - 1) // m * m
    macs = sum(pad(K, 32) * pad(Nc, 64) for K, Nc in gemm) * ns * (ns + 1) // 2
    kms = phase['features'] + phase['contract']
    achieved = 2. * macs * N / (kms * 1e-3) / 1e12
    measured = None
    try: measured = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops')
    except Exception: measured = None
    peak = 2. * measured if measured else 4500.
    nlayers = len(gemm) - len(plan.lowered.blocks)
    layer_flops = 2 * sum(a * b for a, b in gemm[:nlayers])   # FP64 multiply-adds of the hidden layers these GEMMs replace
    fp64_equiv = (flops['contract'] + flops['gram'] + layer_flops) * N / (kms * 1e-3) / 1e12
    return {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TOP/s (int8)', 'frac': achieved / peak, 'traffic': traffic,
            'kernel': 'dlb::ozaki_gemm_kernel (tcgen05.mma kind::i8, TMEM accumulators; {:d} exact 7-bit slices per operand)'.format(ns),
            'gemms_per_step': len(gemm), 'ms_per_step_in_kernel': kms, 'int8_macs_per_eval': macs, 'flops_per_eval_folded': flops,
            'peak_source': ('2 x MEASURED_PEAKS.json bf16_tflops (burst)' if measured else 'nominal dense int8 (no MEASURED_PEAKS.json)'),
            'fp64_equivalent_tflops': fp64_equiv, 'fp64_equivalent_vs_dmma_peak': fp64_equiv / peaks['dmma'],
            'share_of_step': kms / (kms + phase['solve']) if kms else None, 'phase_ms': phase,
            'note': 'time includes the slicing, first-layer and Gram-reduction kernels around the GEMMs'}


def run_workload(ctx, args, workload, N, steps, warmup, main, peaks):
    """One workload: device-resident K steps (+ gather at N > 1 GPUs), sustained leg, e2e leg, per-kernel timing, CPU baselines."""
    import desilike_b200
    torch, rank, world, dev = ctx.torch, ctx.rank, ctx.world, ctx.dev
    t_start = time.perf_counter()
    spec, label = build_spec(workload)
    plan = desilike_b200.Plan(spec, with_theory=False, mlp_slices=6 if workload.endswith('_i8') else 0)
    d, ns = plan.d, plan.ns
    npool = pool_size(N, d)
    pool = [torch.as_tensor(host_batch(spec, N, rank, i), device=dev) for i in range(npool)]
    louts = dict(loglikelihood=torch.empty(N, dtype=torch.float64, device=dev), logprior=torch.empty(N, dtype=torch.float64, device=dev),
                 status=torch.empty(N, dtype=torch.int32, device=dev))
    outs = dict(louts, logposterior=torch.empty(N, dtype=torch.float64, device=dev))
    if ns: outs['x_solved_plan_order'] = torch.empty((N, ns), dtype=torch.float64, device=dev)
    total = world * N
    gather, transport, nbuf = None, None, 4
    gathered, works = None, [None, None]
    if world > 1:
        # The path's only exchange (reference base.py:335, samplers/base.py:200): results of all points to every rank.  Peer-store transport:
        # the finish kernel writes straight into every rank's window; the acquire of step i is enqueued after the evaluation of step i + 1, so
        # the NVLink stores of one step overlap the next step's kernels (4 window buffers make that safe for a consumer of the gathered data).
        from desilike_b200.distributed import PeerGather
        if ctx.comm._want_peer:
            try: gather = PeerGather(ctx.comm, total * (1 + ns), nbuf=nbuf)
            except desilike_b200.DesilikeB200Error as exc:
                log('peer gather unavailable, using NCCL all-gather:', exc)
        transport = 'peer-store (finish kernel -> NVLink st.global into every rank\'s window)' if gather is not None else 'nccl all_gather_into_tensor'
        if gather is None:
            gathered = [torch.empty(total * (1 + ns), dtype=torch.float64, device=dev) for _ in range(2)]

    def step(i):
        if world == 1:
            plan.eval(pool[i % npool], out=outs)
        elif gather is not None:
            plan.eval_gather(pool[i % npool], gather, i % nbuf, rank * N, total, with_solved=ns > 0, out=louts, wait=(i - 1) % nbuf if i > 0 else None)
        else:   # NCCL: in-place all-gather of logposterior (+ x_solved), double buffered so that it overlaps the next evaluation
            b = i % 2
            if works[b] is not None:
                for w in works[b]: w.wait()
            outs['logposterior'] = gathered[b][rank * N:(rank + 1) * N]
            if ns: outs['x_solved_plan_order'] = gathered[b][total + rank * N * ns:total + (rank + 1) * N * ns].view(N, ns)
            plan.eval(pool[i % npool], out=outs)
            works[b] = [ctx.dist.all_gather_into_tensor(gathered[b][:total], outs['logposterior'], async_op=True)]
            if ns: works[b].append(ctx.dist.all_gather_into_tensor(gathered[b][total:], outs['x_solved_plan_order'].view(-1), async_op=True))

    def drain(last):
        if world > 1 and gather is not None and last >= 0: gather.wait(last % nbuf)
        for b in range(2):
            if works[b] is not None:
                for w in works[b]: w.wait()
                works[b] = None

    it = 0
    for _ in range(warmup): step(it); it += 1
    drain(it - 1)
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    launches0 = desilike_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ctx.barrier()
    e0.record()
    first = it
    for _ in range(steps): step(it); it += 1
    drain(it - 1)
    e1.record()
    ctx.barrier()
    launches = desilike_b200.launch_count() - launches0
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / steps
    value = total * steps / (ms_total * 1e-3)
    if gather is not None: gather.check()

    # ---- gather check: every rank's slice of the gathered window of the LAST timed step == a local evaluation of that rank's inputs ----
    gather_check = None
    if world > 1:
        last = it - 1
        window = gather.buffer(last % nbuf).clone() if gather is not None else gathered[last % 2].clone()
        maxdiff, ok = 0., True
        for r in range(world):
            Xr = pool[last % npool] if r == rank else torch.as_tensor(host_batch(spec, N, r, last % npool), device=dev)
            ref = plan.eval(Xr, out=dict(logposterior=torch.empty(N, dtype=torch.float64, device=dev), loglikelihood=louts['loglikelihood'], logprior=louts['logprior'],
                                         **({'x_solved_plan_order': torch.empty((N, ns), dtype=torch.float64, device=dev)} if ns else {})))
            pairs = [(window[r * N:(r + 1) * N], ref['logposterior'])]
            if ns: pairs.append((window[total + r * N * ns:total + (r + 1) * N * ns], ref['x_solved_plan_order'].view(-1)))
            for got, want in pairs:
                same = (got == want) | (torch.isnan(got) & torch.isnan(want))
                ok = ok and bool(same.all())
                fin = torch.isfinite(got) & torch.isfinite(want)
                if bool(fin.any()): maxdiff = max(maxdiff, float((got[fin] - want[fin]).abs().max()))
        gather_check = {'ok': ok, 'ranks_checked': world, 'max_abs_diff': maxdiff, 'what': 'gathered logposterior' + (' + x_solved' if ns else '') + ' of the last timed step vs a local evaluation of every rank\'s inputs (bitwise)'}
        flag = ctx.max_over_ranks(0. if ok else 1.)
        if flag: raise RuntimeError('gather check FAILED on some rank: {}'.format(gather_check))

    # ---- sustained leg: the same steps back to back for >= args.sustained seconds (same count on every rank), clocks sampled throughout ----
    sustained = None
    if args.sustained > 0:
        nsus = int(max(steps, np.ceil((args.sustained if main else min(args.sustained, 0.5)) * 1e3 / ms_per_step)))
        ctx.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(nsus): step(it); it += 1
        drain(it - 1)
        s1.record()
        ctx.barrier()
        ms_sus = ctx.max_over_ranks(s0.elapsed_time(s1))
        sustained = {'value': total * nsus / (ms_sus * 1e-3), 'unit': UNIT, 'steps': nsus, 'seconds': ms_sus * 1e-3, 'ms_per_step': ms_sus / nsus}
    clocks = sampler.result()
    clocks['window'] = 'timed K steps + sustained leg ({:.2f} s under load)'.format((ms_total + (sustained['seconds'] * 1e3 if sustained else 0.)) * 1e-3)

    # ---- e2e: public API with pinned host buffers, H2D + D2H inside the timed region ----
    hosts = [desilike_b200.pinned_empty((N, d)) for _ in range(4)]
    for i, h in enumerate(hosts): h[...] = pool[i % npool].cpu().numpy()
    hout = dict(logposterior=desilike_b200.pinned_empty(N), loglikelihood=desilike_b200.pinned_empty(N), logprior=desilike_b200.pinned_empty(N),
                status=desilike_b200.pinned_empty(N, np.int32))
    if ns: hout['x_solved_plan_order'] = desilike_b200.pinned_empty((N, ns))
    e2e_steps = max(3, min(steps, 200 if ms_per_step < 1. else 40))
    if world == 1:
        def e2e_step(i): plan.eval(hosts[i % 4], out=hout)
        api = 'desilike_b200.Plan.eval(pinned host ndarray) -> dlb_eval_batch'
        h2d, d2h = N * d * 8, N * (3 * 8 + 4) + N * ns * 8
    elif gather is not None:
        # what a host-side consumer of a sharded evaluation reads back (reference samplers/base.py:196-200: logposterior of ALL points on every
        # rank; the derived values -- here the solved parameters -- of its own points): gathered logposterior + own x_solved slice + local outputs
        hgat = torch.empty(total, dtype=torch.float64, pin_memory=True)
        lhost = {k: hout[k] for k in ('loglikelihood', 'logprior', 'status') + (('x_solved_plan_order',) if ns else ())}   # own points: pipelined copy-out

        def e2e_step(i):   # host parameters in -> finish kernel stores into every rank's window -> results back to the host
            b = i % nbuf
            plan.eval_gather(hosts[i % 4], gather, b, rank * N, total, with_solved=ns > 0, out=lhost)
            gather.wait(b)
            window = gather.buffer(b, total * (1 + ns))
            hgat.copy_(window[:total], non_blocking=True)
            torch.cuda.synchronize()
        api = 'desilike_b200.Plan.eval_gather(pinned host ndarray, own loglikelihood / logprior / status / x_solved copied out per sub-batch) -> dlb_eval_batch_gather; PeerGather.wait; D2H of the gathered logposterior'
        h2d, d2h = N * d * 8, N * (2 * 8 + 4) + total * 8 + N * ns * 8
    else:
        def e2e_step(i):
            plan.eval(hosts[i % 4], out=hout)
            gathered[0][rank * N:(rank + 1) * N].copy_(torch.from_numpy(hout['logposterior']), non_blocking=True)
            ctx.dist.all_gather_into_tensor(gathered[0][:total], gathered[0][rank * N:(rank + 1) * N])
        api = 'desilike_b200.Plan.eval(pinned host ndarray) + NCCL all-gather of logposterior'
        h2d, d2h = N * d * 8 + N * 8, N * (3 * 8 + 4) + N * ns * 8
    for i in range(3): e2e_step(i)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps): e2e_step(i)
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {'value': total * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'api': api}

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    plan.set_timing(True)
    tms = {'features': [], 'contract': [], 'solve': [], 'fftlog': []}
    for i in range(max(5, min(steps, 30))):
        plan.eval(pool[i % npool], out=outs if world == 1 else dict(louts, logposterior=torch.empty(N, dtype=torch.float64, device=dev)))
        for k, v in plan.last_kernel_ms().items(): tms[k].append(v)
    plan.set_timing(False)
    phase = {k: float(np.mean(v[2:])) for k, v in tms.items()}
    roofline = roofline_for(workload, plan, phase, N, value / world, peaks)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(label, N, world, d), 'clocks': clocks, 'sustained': sustained, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline}
    if world > 1:
        per_peer = N * (1 + ns) * 8
        line['exchange'] = {'transport': transport, 'nvlink_bytes_per_step_per_gpu': per_peer * (world - 1), 'bytes_per_step_per_peer': per_peer,
                            'collective_kernels_per_step': 0 if gather is not None else (2 if ns else 1),
                            'what': 'logposterior [N] + x_solved [N, ns] of this rank stored into each of the {:d} peers'.format(world - 1)}
        line['gather_check'] = gather_check
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_seconds = args.cpu_seconds if main else min(args.cpu_seconds, 3.)
        Xh = hosts[0]
        cpu_value, cpu_done = cpu_port(spec, Xh, seconds=cpu_seconds, max_points=2048 if spec.get('solved') else 32768)
        line['cpu_baseline'] = {'value': cpu_value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': '{:d} points of the same workload, oracle/pipeline.py (vectorised NumPy, threaded BLAS), {:.0f} s'.format(cpu_done, cpu_seconds)}
        if workload in ('c2', 'c4', 'c5', 'c5solved', 'c4wide') and not plan.lowered.two_stage and not plan.lowered.split_dense:
            try:
                fv, fdone = cpu_folded_rate(plan.lowered, Xh, min(cpu_seconds, 4.))
                line['cpu_baseline']['folded'] = {'value': fv, 'unit': UNIT, 'cores': os.cpu_count(),
                                                  'what': 'the plan-time-folded chain the GPU executes (lowered term matrices), vectorised NumPy, {:d} points'.format(fdone)}
            except Exception as exc:   # noqa: BLE001
                log('folded CPU leg failed:', exc)
        if main: line['cpu_baseline']['reference_loop_recorded'] = recorded_reference_loop()
    if gather is not None:
        ctx.barrier()
        gather.close()
    plan.close()
    del pool, outs, louts, hosts, hout
    torch.cuda.empty_cache()
    log('{} done in {:.1f} s: value {:.4g} e2e {:.4g} frac {:.3f}'.format(workload, time.perf_counter() - t_start, value, e2e['value'], roofline['frac']))
    return line


def run_ensemble(ctx, args, steps, warmup, peaks):
    """BASELINE.json configs[4]: one ensemble-sampler iteration (stretch move, red-blue halves) for 2^20 walkers of the 5-tracer
    Taylor-emulated likelihood, walkers sharded over the GPUs (strong scaling); every iteration the log-posteriors of all walkers reach
    every rank (peer stores over NVLink, or an NCCL all-gather).  Walkers, proposals and log-posteriors are device resident
    (dlb_ensemble_step); `value` = log-posterior evaluations / s."""
    import desilike_b200
    configs = workload_module()
    torch, rank, world, dev = ctx.torch, ctx.rank, ctx.world, ctx.dev
    spec = configs.config_c5(coefficients='random')
    plan = desilike_b200.Plan(spec, with_theory=False)
    total = 1 << 20
    n = total // world
    X = configs.sample_points(spec, n, seed=500 + rank)
    X = X.mean(axis=0) + 0.05 * (X - X.mean(axis=0))
    es = desilike_b200.EnsembleSampler(plan, n, seed=2026, rank_offset=rank * (n // 2))
    es.set_start(X)
    gather, gathered, nbuf = None, None, 2
    if world > 1:
        from desilike_b200.distributed import PeerGather
        if ctx.comm._want_peer:
            try: gather = PeerGather(ctx.comm, total, nbuf=nbuf)
            except desilike_b200.DesilikeB200Error as exc: log('peer gather unavailable:', exc)
        if gather is None: gathered = torch.empty((world, n), dtype=torch.float64, device=dev)
    state = {'i': 0}

    def step():
        es.step()
        if gather is not None:   # immediate acquire: the accept / reject of the next iteration may read the gathered values
            b = state['i'] % nbuf
            gather.publish(b, es.logposterior, rank * n)
            gather.wait(b)
        elif world > 1: ctx.dist.all_gather_into_tensor(gathered, es.logposterior)
        state['i'] += 1

    for _ in range(max(warmup, 3)): step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    launches0 = desilike_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ctx.barrier()
    e0.record()
    for _ in range(steps): step()
    e1.record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = desilike_b200.launch_count() - launches0
    clocks = sampler.result()
    if gather is not None: gather.check()
    value = total * steps / (ms_total * 1e-3)
    # e2e: the same iteration driven from HOST walkers each step (H2D of the walkers + D2H of positions and log-posteriors)
    hX = desilike_b200.pinned_empty((n, plan.d)); hX[...] = es.walkers.cpu().numpy()
    hlp = desilike_b200.pinned_empty(n)
    e2e_steps = max(3, min(steps, 20))
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        es.walkers.copy_(torch.from_numpy(hX), non_blocking=True)
        step()
        torch.from_numpy(hX).copy_(es.walkers, non_blocking=True)
        torch.from_numpy(hlp).copy_(es.logposterior, non_blocking=True)
        torch.cuda.synchronize()
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    flops = plan.lowered.flops_per_point()
    peak = peaks['dmma']
    achieved = (flops['contract'] + flops['gram']) * value / world / 1e12
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': max(warmup, 3), 'ms_per_step': ms_total / steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C5 ensemble step: 2^20 walkers, stretch move (2 half-ensemble batches / iteration), 5-tracer Taylor likelihood (n=2010, d=14), device-resident walkers',
                       'walkers_per_gpu': n, 'parallelism': 'walkers sharded over {:d} GPU(s); logposterior of all walkers delivered to every rank per iteration ({})'.format(
                           world, 'peer stores over NVLink' if gather is not None else ('NCCL all-gather' if world > 1 else 'single GPU')),
                       'l2_policy': 'proposals ({:.0f} MB per half) are regenerated every iteration; inputs exceed L2 over an iteration'.format(n // 2 * plan.d * 8 / 2**20)},
            'clocks': clocks, 'acceptance_fraction': es.acceptance_fraction,
            'e2e': {'value': total * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': n * plan.d * 8, 'd2h_bytes_per_step': n * (plan.d + 1) * 8, 'steps': e2e_steps,
                    'api': 'desilike_b200.EnsembleSampler.step with host walkers copied in and out every iteration'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                         'kernel': 'dlb::contract_kernel inside the whole iteration (propose + likelihood + accept)', 'flops_per_eval_folded': flops,
                         'flops_per_eval_executed': plan.executed_flops(), 'frac_executed': plan.executed_flops() * value / world / 1e12 / peak,
                         'peak_source': 'dlb_measure_fp64_peak(DMMA.8x8x4 register loop) measured in this run'}}
    if gather is not None:
        ctx.barrier()
        gather.close()
    plan.close()
    torch.cuda.empty_cache()
    return line


def compact(line):
    """Extra-workload entry carried inside the main JSON line."""
    r = line.get('roofline') or {}
    out = {'workload': line['config']['workload'], 'batch_per_gpu': line['config'].get('batch_per_gpu', line['config'].get('walkers_per_gpu')), 'value': line['value'],
           'unit': line['unit'], 'ms_per_step': line['ms_per_step'], 'steps': line['steps'], 'scaling': line['scaling'],
           'e2e': {k: line['e2e'][k] for k in ('value', 'h2d_bytes_per_step', 'd2h_bytes_per_step')},
           'roofline': {k: r.get(k) for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'frac_executed', 'traffic', 'kernel', 'ms_per_step_in_kernel', 'flops_per_eval_kernel', 'phase_ms')},
           'gpu_launches': line['gpu_launches']}
    if line.get('sustained'): out['sustained'] = {k: line['sustained'][k] for k in ('value', 'seconds')}
    if line.get('cpu_baseline'): out['cpu_baseline'] = line['cpu_baseline']
    if line.get('exchange'): out['exchange'] = line['exchange']
    if line.get('gather_check'): out['gather_check'] = line['gather_check']
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=100)
    parser.add_argument('--warmup', type=int, default=5)
    parser.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    parser.add_argument('--workload', default=DEFAULT_WORKLOAD)
    parser.add_argument('--batch', type=int, default=None)
    parser.add_argument('--cpu-seconds', type=float, default=10.)
    parser.add_argument('--no-cpu', action='store_true')
    parser.add_argument('--extras', default='auto', help="'auto' (all other workloads at 1 GPU; c2 + c5step at N > 1 GPUs), 'none', or a comma list")
    parser.add_argument('--sustained', type=float, default=2., help='seconds of the back-to-back sustained leg (0: off)')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.batch is None:
        args.batch = DEFAULT_BATCH if args.workload == DEFAULT_WORKLOAD else EXTRA_BATCH.get(args.workload, 65536)

    global _IMPL
    _IMPL = args.impl
    configs = workload_module()
    if args.impl == 'reference':
        spec, label = build_spec('c5' if args.workload == 'c5step' else args.workload)
        X = configs.sample_points(spec, min(args.batch, 32768), seed=100)
        return run_reference(args, spec, label, X)

    ctx = Context()
    peaks = fp64_peaks(ctx)
    if args.workload == 'c5step':
        line = run_ensemble(ctx, args, args.steps, args.warmup, peaks)
    else:
        line = run_workload(ctx, args, args.workload, args.batch, args.steps, args.warmup, True, peaks)
    line['fp64_peaks_tflops'] = peaks
    if args.extras == 'auto':
        extras = [w for w in ('c2', 'c5', 'c5solved', 'c1', 'c3', 'c3fft', 'c4', 'c4dense', 'c4wide', 'c4wide_i8', 'c5step') if w != args.workload] if ctx.world == 1 else \
                 [w for w in ('c2', 'c5step') if w != args.workload]
        if args.workload != DEFAULT_WORKLOAD: extras = []
    elif args.extras == 'none': extras = []
    else: extras = [w for w in args.extras.split(',') if w]
    line['extra'] = []
    for w in extras:
        try:
            if w == 'c5step':
                ex = run_ensemble(ctx, args, max(3, min(args.steps, 10)), 3, peaks)
            else:
                ex = run_workload(ctx, args, w, EXTRA_BATCH.get(w, 65536), max(5, min(args.steps, 10 if w in ('c4wide', 'c4wide_i8', 'c4dense') else 100)), 3, False, peaks)
            line['extra'].append(compact(ex))
        except Exception as exc:   # noqa: BLE001 -- an extra must not take the main line down; it is reported as failed
            if ctx.world > 1: raise
            log('extra workload', w, 'failed:', repr(exc))
            line['extra'].append({'workload': w, 'failed': repr(exc)})
    if ctx.rank == 0:
        emit(line)
    if ctx.world > 1:
        ctx.comm.close()
        ctx.dist.destroy_process_group()


if __name__ == '__main__':
    main()
