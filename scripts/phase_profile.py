"""Per-phase cycle breakdown of the contraction kernel (needs the -DDLB_PROFILE build: lib/libdesilike_b200_prof.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['DESILIKE_B200_LIB'] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'desilike_b200', 'lib', 'libdesilike_b200_prof.so')
os.environ['DLB_PROFILE_PRINT'] = '1'
import torch
import bench, desilike_b200
from oracle import configs
for w in sys.argv[1:]:
    spec, label = bench.build_spec(w)
    plan = desilike_b200.Plan(spec, with_theory=False)
    X = torch.as_tensor(configs.sample_points(spec, 65536, seed=1), device='cuda')
    print(w, flush=True)
    for i in range(3): plan.eval(X)
    torch.cuda.synchronize()
