"""DMMA dependent-chain latency: rate vs independent accumulators per warp and warps per SM (dlb_measure_fp64_peak kinds 2000+)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import desilike_b200
for warps in (4, 8, 16):
    print('warps/SM', warps, ' '.join('ilp{}={:.2f}'.format(ilp, desilike_b200.measure_fp64_peak(2000 + 100 * ilp + warps)) for ilp in (1, 2, 4, 6, 10, 14)))
