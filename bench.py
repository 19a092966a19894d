#!/usr/bin/env python
"""Benchmark of the batched full-shape likelihood path (BASELINE.json metric: logL evaluations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c5|c5solved|c4] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of ``--batch`` (default 65 536) synthetic parameter points.
Default workload = BASELINE.json configs[1] ("TaylorEmulator (order 4) of full-shape P_ell(k) + window matrix + chi2,
65k params on 1xB200").  Prints ONE JSON line (rank 0).

  value     device-resident throughput: inputs already in HBM, K steps bracketed by barrier + synchronize, CUDA events,
            max over ranks.  Inputs rotate through a pool of distinct batches larger than L2 (no L2 flush needed).
  e2e       same metric through the public host API (Plan.eval on pinned HOST buffers): H2D of the parameters and D2H of
            logposterior / loglikelihood / logprior / status inside the timed region.
  roofline  dominant kernel (FP64 DMMA contraction): algorithmic FLOPs of the folded chain per launch / CUDA-event time
            of those launches, against the FP64 DMMA peak measured in this run (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  the NumPy oracle port (threaded BLAS) on a bounded sample of the same workload on this box's host cores.

Synthetic inputs: the workload definitions (spec constants, fiducial data vector, covariance) of SURVEY.md section 8d live in
oracle/configs.py, shared with the tests; bench.py imports them to BUILD its inputs before any timing.  Every timed region of the default
arm executes desilike_b200 (the C-ABI library) only; oracle/pipeline.py runs in the cpu_baseline leg and in ``--impl reference``.

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


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'logL evals/sec (batched, emulated FS+window+chi2)'
# algorithmic FLOP per evaluation of the UN-collapsed chain
UNCOLLAPSED_FLOPS = {'c1': 0.1e6, 'c2': 0.7e6, 'c3': 1.3e6, 'c3fft': 1.3e6, 'c4': 20e6, 'c4wide': 47e6, 'c5': 8e6, 'c5solved': 19e6}
UNIT = 'evals/s'


def build_spec(workload):
    from oracle import configs
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
        return configs.config_c4(), 'C4: 5 tracers x [MLP 4-64-64-64 pt emulator -> EFT-like combine -> window [402,1050]] -> chi2 + marginalisation (ns=35)'
    if workload == 'c4wide':
        return configs.config_c4(hidden=(512, 512, 512, 512)), 'C4 stress variant: 5 tracers x [MLP 4-512-512-512-512 pt emulator -> EFT-like combine -> window [402,1050]] -> chi2 + marginalisation (ns=35)'
    raise ValueError(workload)


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


def run_reference(args, spec, label, X):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import pipeline
    pipeline.evaluate(spec, X[:256])
    t0 = time.perf_counter()
    pipeline.evaluate(spec, X[:1024])
    speed = 1024 / (time.perf_counter() - t0)
    # bounded sample per step: the whole --steps K run stays within ~1-2 minutes of CPU time
    sample = int(min(len(X), max(256, speed * 60. / max(args.steps, 1))))
    for _ in range(min(args.warmup, 2)): pipeline.evaluate(spec, X[:sample])
    t0 = time.perf_counter()
    for _ in range(args.steps): pipeline.evaluate(spec, X[:sample])
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cores = os.cpu_count()
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': label, 'batch': sample, 'note': 'CPU: each step is a bounded sample of the GPU batch'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '{:d} points/step x {:d} steps, oracle/pipeline.py (vectorised NumPy, threaded BLAS)'.format(sample, args.steps)},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    emit(line)


def run_ensemble(args):
    """BASELINE.json configs[4]: one ensemble-sampler iteration (stretch move, red-blue halves) for 2^20 walkers of the 5-tracer
    Taylor-emulated likelihood, walkers sharded over the GPUs (strong scaling), NCCL all-gather of the log-posteriors per step.
    Walkers, proposals and log-posteriors are device resident (dlb_ensemble_step); `value` = log-posterior evaluations / s."""
    import torch
    import desilike_b200
    from oracle import configs
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    spec = configs.config_c5(coefficients='random')
    plan = desilike_b200.Plan(spec, with_theory=False)
    total = 1 << 20
    n = total // world
    X = configs.sample_points(spec, n, seed=500 + rank)
    X = X.mean(axis=0) + 0.05 * (X - X.mean(axis=0))
    es = desilike_b200.EnsembleSampler(plan, n, seed=2026, rank_offset=rank * (n // 2))
    es.set_start(X)
    gathered = torch.empty((world, n), dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        es.step()
        if world > 1: dist.all_gather_into_tensor(gathered, es.logposterior)

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)): step()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = desilike_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps): step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = desilike_b200.launch_count() - launches0
    clocks = sampler.result()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = total * args.steps / (ms_total * 1e-3)
    # e2e: the same iteration driven from HOST walkers each step (H2D of the walkers + D2H of positions and log-posteriors)
    hX = desilike_b200.pinned_empty((n, plan.d)); hX[...] = es.walkers.cpu().numpy()
    hlp = desilike_b200.pinned_empty(n)
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        es.walkers.copy_(torch.from_numpy(hX), non_blocking=True)
        step()
        torch.from_numpy(hX).copy_(es.walkers, non_blocking=True)
        torch.from_numpy(hlp).copy_(es.logposterior, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    flops = plan.lowered.flops_per_point()
    peak = desilike_b200.measure_fp64_peak('dmma')
    achieved = (flops['contract'] + flops['gram']) * value / world / 1e12
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C5 ensemble step: 2^20 walkers, stretch move (2 half-ensemble batches / iteration), 5-tracer Taylor likelihood (n=2010, d=14), device-resident walkers',
                       'walkers_per_gpu': n, 'parallelism': 'walkers sharded over {:d} GPU(s); all-gather of logposterior per iteration'.format(world),
                       'l2_policy': 'proposals ({:.0f} MB per half) are regenerated every iteration; inputs exceed L2 over an iteration'.format(n // 2 * plan.d * 8 / 2**20)},
            'clocks': clocks, 'acceptance_fraction': es.acceptance_fraction,
            'e2e': {'value': total * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': n * plan.d * 8, 'd2h_bytes_per_step': n * (plan.d + 1) * 8, 'steps': e2e_steps,
                    'api': 'desilike_b200.EnsembleSampler.step with host walkers copied in and out every iteration'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': None,
                         'kernel': 'dlb::contract_kernel inside the whole iteration (propose + likelihood + accept)', 'flops_per_eval_folded': flops,
                         'flops_per_eval_executed': plan.executed_flops(), 'frac_executed': plan.executed_flops() * value / world / 1e12 / peak,
                         'peak_source': 'dlb_measure_fp64_peak(DMMA.8x8x4 register loop) measured in this run'}}
    if rank == 0: emit(line)
    if world > 1: dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=400)
    parser.add_argument('--warmup', type=int, default=10)
    parser.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    parser.add_argument('--workload', default='c2')
    parser.add_argument('--batch', type=int, default=65536)
    parser.add_argument('--cpu-seconds', type=float, default=10.)
    parser.add_argument('--no-cpu', action='store_true')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    from oracle import configs
    if args.workload == 'c5step' and args.impl == 'ours':
        return run_ensemble(args)
    spec, label = build_spec('c5' if args.workload == 'c5step' else args.workload)
    N = args.batch
    if args.impl == 'reference':
        X = configs.sample_points(spec, N, seed=100)
        return run_reference(args, spec, label, X)

    import torch
    import desilike_b200
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    plan = desilike_b200.Plan(spec, with_theory=False)
    d = plan.d

    # rotating pool of distinct input batches, larger than L2 (126 MB)
    l2_bytes = 126 * 2**20
    npool = max(2, int(np.ceil(1.25 * l2_bytes / (N * d * 8))))
    pool = [torch.as_tensor(configs.sample_points(spec, N, seed=1000 + 97 * rank + i), device=dev) for i in range(min(npool, 8))]
    while len(pool) < npool:   # cheap distinct batches: jitter within the prior box keeps the workload identical
        base = pool[len(pool) % 8]
        pool.append((base * (1. + 1e-6 * len(pool))).contiguous())
    outs = dict(logposterior=torch.empty(N, dtype=torch.float64, device=dev), loglikelihood=torch.empty(N, dtype=torch.float64, device=dev),
                logprior=torch.empty(N, dtype=torch.float64, device=dev), status=torch.empty(N, dtype=torch.int32, device=dev))
    if plan.ns: outs['x_solved_plan_order'] = torch.empty((N, plan.ns), dtype=torch.float64, device=dev)
    # The path's only exchange: all-gather of the log-posteriors (reference base.py:335, samplers/base.py:200).  The finish kernel
    # writes straight into this rank's slice of the gather buffer (in-place collective); two buffers alternate so that the NCCL
    # all-gather of step i (its own stream) overlaps the evaluation of step i + 1 -- no rank needs the other ranks' values to go on.
    gathered = [torch.empty(world * N, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    works = [None, None]

    def step(i):
        if world > 1:
            b = i % 2
            if works[b] is not None: works[b].wait()
            outs['logposterior'] = gathered[b][rank * N:(rank + 1) * N]
            plan.eval(pool[i % npool], out=outs)
            works[b] = dist.all_gather_into_tensor(gathered[b], outs['logposterior'], async_op=True)
        else:
            plan.eval(pool[i % npool], out=outs)

    def barrier():
        if world > 1:
            for b in range(2):
                if works[b] is not None: works[b].wait(); works[b] = None
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup): step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = desilike_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for i in range(args.steps): step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = desilike_b200.launch_count() - launches0
    if len(sampler.samples) < 5:   # timed region too short for NVML sampling: keep sampling over an untimed replay of the same steps
        t_end = time.perf_counter() + 0.5
        i = 0
        while time.perf_counter() < t_end:   # no collective here: the number of replayed steps is time based, hence rank dependent
            plan.eval(pool[i % npool], out=outs); i += 1
            if i % 32 == 0: torch.cuda.synchronize()
        torch.cuda.synchronize()
    clocks = sampler.result()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * N * args.steps / (ms_total * 1e-3)

    # ---- e2e: public API with pinned host buffers, H2D + D2H inside the timed region ----
    hosts = [desilike_b200.pinned_empty((N, d)) for _ in range(4)]
    for i, h in enumerate(hosts): h[...] = pool[i].cpu().numpy()
    hout = dict(logposterior=desilike_b200.pinned_empty(N), loglikelihood=desilike_b200.pinned_empty(N), logprior=desilike_b200.pinned_empty(N),
                status=desilike_b200.pinned_empty(N, np.int32))
    if plan.ns: hout['x_solved_plan_order'] = desilike_b200.pinned_empty((N, plan.ns))
    e2e_steps = max(3, min(args.steps, 200))
    for i in range(3): plan.eval(hosts[i % 4], out=hout)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        plan.eval(hosts[i % 4], out=hout)
        if world > 1:   # the host result goes back to the device-side gather buffer: H2D of this rank's slice + all-gather
            gathered[0][rank * N:(rank + 1) * N].copy_(torch.from_numpy(hout['logposterior']), non_blocking=True)
            dist.all_gather_into_tensor(gathered[0], gathered[0][rank * N:(rank + 1) * N])
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * N * e2e_steps / dt
    h2d = N * d * 8
    d2h = N * (3 * 8 + 4) + (N * plan.ns * 8 if plan.ns else 0)

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    flops = plan.lowered.flops_per_point()
    plan.set_timing(True)
    tms = {'features': [], 'contract': [], 'solve': [], 'fftlog': []}
    for i in range(max(5, min(args.steps, 30))):
        plan.eval(pool[i % npool], out=outs)
        for k, v in plan.last_kernel_ms().items(): tms[k].append(v)
    plan.set_timing(False)
    phase = {k: float(np.mean(v[2:])) for k, v in tms.items()}
    step_ms = phase['features'] + phase['contract'] + phase['solve']
    peak_dmma = desilike_b200.measure_fp64_peak('dmma')
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'traffic_{}.json'.format(args.workload))
    if os.path.exists(prof):
        try: traffic = json.load(open(prof)).get('dram_bytes_per_launch')
        except Exception: traffic = None
    uncollapsed = UNCOLLAPSED_FLOPS.get(args.workload)
    if args.workload == 'c3':
        # dominant kernel: damped-BAO P_ell(kin) (one warp per point; FP64 FMA pipe).  Algorithmic work (SURVEY.md section 8a row a12,
        # 8d "BAO P(k)"): nk * nmu nodes x (3 cubic interpolations ~ 20 FLOP each + damping exp ~ 30 + AP / Kaiser / FoG factors ~ 40)
        # + the Legendre sums 2 nell nk nmu, with nk = 300, nmu = 10, nell = 2
        nk_, nmu_, nell_ = 300, 10, 2
        bao_flops = nk_ * nmu_ * (3 * 20 + 30 + 40) + 2 * nell_ * nk_ * nmu_
        peak = desilike_b200.measure_fp64_peak('dfma')
        achieved = bao_flops * N / (phase['features'] * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'kernel': 'dlb::bao_kernel (FP64 FMA pipe; AP + wiggle spline + Gauss-Legendre quadrature, one warp per point)', 'launches_per_step': 1,
                    'ms_per_step_in_kernel': phase['features'], 'flops_per_eval': {'bao': bao_flops, 'contract_folded': flops},
                    'peak_source': 'dlb_measure_fp64_peak(DFMA register loop) measured in this run; MEASURED_PEAKS.json has no FP64 entry',
                    'share_of_step': phase['features'] / step_ms, 'phase_ms': phase}
    elif args.workload == 'c3fft':
        # dominant kernel: FFTLog (one CTA per point and multipole; FP64 FMA pipe + shared-memory passes).  Algorithmic work =
        # n_ell (2 * 2.5 Np log2 Np + 6 Np) FLOP per evaluation with Np = 4096.
        nell, Np = 2, 4096
        fft_flops = nell * (2 * 2.5 * Np * np.log2(Np) + 6 * Np)
        peak = desilike_b200.measure_fp64_peak('dfma')
        achieved = fft_flops * N / (phase['fftlog'] * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'kernel': 'dlb::fftlog_kernel (FP64 FMA pipe, radix 16x16x8 passes in shared memory)', 'launches_per_step': 1,
                    'ms_per_step_in_kernel': phase['fftlog'], 'flops_per_eval': {'fftlog': fft_flops},
                    'peak_source': 'dlb_measure_fp64_peak(DFMA register loop) measured in this run (no tensor-core FP64 FFT path); MEASURED_PEAKS.json has no FP64 entry',
                    'share_of_step': phase['fftlog'] / step_ms, 'phase_ms': phase}
    elif args.workload == 'c1':
        # dominant kernel: Kaiser + ShapeFit + AP (one warp per point).  Algorithmic work = SURVEY.md section 8d:
        # Nkin * nmu * 60 + 500 * (exp, tanh) ~ 0.05 MFLOP per evaluation (special functions at ~30 FLOP each).
        kaiser_flops = 36 * 8 * 60 + 500 * 2 * 30
        peak = desilike_b200.measure_fp64_peak('dfma')
        achieved = kaiser_flops * N / (phase['features'] * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
                    'kernel': 'dlb::kaiser_kernel (FP64 FMA pipe; spline + Gauss-Legendre quadrature, one warp per point)', 'launches_per_step': 1,
                    'ms_per_step_in_kernel': phase['features'], 'flops_per_eval': {'kaiser': kaiser_flops},
                    'peak_source': 'dlb_measure_fp64_peak(DFMA register loop) measured in this run; MEASURED_PEAKS.json has no FP64 entry',
                    'share_of_step': phase['features'] / step_ms, 'phase_ms': phase}
    else:
        contract_ms = phase['contract']
        achieved = (flops['contract'] + flops['gram']) * N / (contract_ms * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_dmma, 'unit': 'TFLOP/s', 'frac': achieved / peak_dmma, 'traffic': traffic,
                    'kernel': 'dlb::contract_kernel (FP64 DMMA.8x8x4, persistent, one launch per likelihood)', 'launches_per_step': 1,
                    'ms_per_step_in_kernel': contract_ms, 'flops_per_eval_folded': flops,
                    'peak_source': 'dlb_measure_fp64_peak(DMMA.8x8x4 register loop) measured in this run; MEASURED_PEAKS.json has no FP64 entry',
                    'share_of_step': contract_ms / step_ms, 'phase_ms': phase}
    if args.workload not in ('c1', 'c3', 'c3fft'):   # what the kernel multiplies (padding inside tiles included) next to the algorithmic count above
        roofline['flops_per_eval_executed'] = plan.executed_flops()
        roofline['achieved_executed'] = plan.executed_flops() * N / (phase['contract'] * 1e-3) / 1e12
        roofline['frac_executed'] = roofline['achieved_executed'] / peak_dmma
    if uncollapsed:   # the reference's un-collapsed chain: what the same evaluations cost without plan-time folding
        roofline['flops_per_eval_uncollapsed'] = uncollapsed
        roofline['achieved_uncollapsed_equivalent'] = uncollapsed * value / 1e12

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': label, 'batch_per_gpu': N, 'parallelism': 'points sharded over {:d} GPU(s); all-gather of logposterior'.format(world) if world > 1 else 'single GPU',
                       'l2_policy': 'inputs rotate through a pool of {:d} distinct batches ({:.0f} MB > 126 MB L2)'.format(npool, npool * N * d * 8 / 2**20)},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                    'api': 'desilike_b200.Plan.eval(pinned host ndarray) -> dlb_eval_batch'},
            'gpu_launches': int(launches), 'roofline': roofline}
    if rank == 0 and world == 1 and not args.no_cpu:
        Xh = hosts[0]
        cpu_value, cpu_done = cpu_port(spec, Xh, seconds=args.cpu_seconds, max_points=2048 if spec.get('solved') else 32768)
        line['cpu_baseline'] = {'value': cpu_value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': '{:d} points of the same workload, oracle/pipeline.py (vectorised NumPy, threaded BLAS), {:.0f} s'.format(cpu_done, args.cpu_seconds)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
