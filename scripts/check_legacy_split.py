"""Non-resident plan (direct Kaiser generator), host batch >= 131 072 points: the equal-split pipeline gives the device-path results bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import desilike_b200
from desilike_b200 import workloads as configs
spec = configs.config_kaiser_window(nkin=200, nbins=8, eftlike=True, solved=True)   # 1202 feature columns: too wide for a resident A tile
plan = desilike_b200.Plan(spec)
X = configs.sample_points(spec, 2 * 65536 + 1234, seed=3)
dev = plan.eval(torch.as_tensor(X, device='cuda')); torch.cuda.synchronize()
host = plan.eval(X)
print('legacy split ok', all(np.array_equal(host[k], dev[k].cpu().numpy()) for k in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']))
