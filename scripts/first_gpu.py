import time, json
import numpy as np
import torch
import desilike_b200
from oracle import configs, pipeline
print('fp64 peaks', desilike_b200.measure_fp64_peak('dmma'), desilike_b200.measure_fp64_peak('dfma'))
for name, spec, N in [('c2', configs.config_c2(coefficients='random'), 65536), ('c5', configs.config_c5(coefficients='random'), 65536),
                      ('c5solved', configs.config_c5(coefficients='random', solved=True), 65536)]:
    plan = desilike_b200.Plan(spec)
    X = configs.sample_points(spec, N, seed=1)
    Xd = torch.as_tensor(X, device='cuda')
    plan.set_timing(True)
    for i in range(3): out = plan.eval(Xd)
    torch.cuda.synchronize()
    print(name, plan.last_kernel_ms(), plan.lowered.flops_per_point())
    plan.set_timing(False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): out = plan.eval(Xd, return_solved=False, return_status=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, 'device-resident ms/step', ms, 'evals/s', N / ms * 1e3)
    t = time.time()
    for i in range(5): o = plan.eval(X)
    dt = (time.time() - t) / 5
    print(name, 'host e2e ms', dt * 1e3, 'evals/s', N / dt)
