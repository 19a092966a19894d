"""Time the REFERENCE's own serial batch loop -- unmodified desilike classes under the stub third-party modules, driven by
`desilike.base.vmap(likelihood, backend=None)` (reference base.py:232-273) -- in the BUILD CONTAINER, 1 core, and record the result in
profiles/reference_loop.json.  /root/reference does not exist on the GPU box, so bench.py ships these numbers as a recorded constant
(BASELINE.md section 4 item 1).  Usage: python scripts/time_reference_loop.py"""
import json
import os
import platform
import sys
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import configs, ref_harness, pipeline  # noqa: E402


def rate(fn, n, seconds=8.):
    fn(min(n, 16))
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn(n); done += n
    return done / (time.perf_counter() - t0)


def main():
    ref_harness.setup()
    out = {'where': 'build container ({}; {} logical cores; 1 core used, OMP_NUM_THREADS=1)'.format(platform.processor() or platform.machine(), os.cpu_count()),
           'what': 'unmodified reference classes under oracle/stubs, desilike.base.vmap(likelihood, backend=None) serial loop (reference base.py:239-273)',
           'script': 'scripts/time_reference_loop.py', 'unit': 'evals/s', 'cases': {}}
    cases = [('c1_direct_kaiser_shapefit_n108', configs.config_c1(), dict()),
             ('kaiser_window_n36_nkin60', configs.config_kaiser_window(eftlike=False), dict()),
             ('eftlike_window_7_marginalised', configs.config_kaiser_window(eftlike=True, solved=True), dict(eftlike=True, solved_mode='.marg'))]
    for name, spec, kw in cases:
        like = ref_harness.reference_kaiser_likelihood(spec, **kw)
        if kw.get('solved_mode'): ref_harness.install_jacfwd_stub()
        X = configs.sample_points(spec, 256, seed=1)
        r_ref = rate(lambda n: ref_harness.run_vmap(like, spec, X[:n], return_theory=False), 64 if kw.get('solved_mode') else 256)
        r_port = rate(lambda n: pipeline.evaluate(spec, X[:n]), 256)
        out['cases'][name] = {'reference_loop': r_ref, 'oracle_port_same_points_1core': r_port}
        print(name, r_ref, r_port)
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'reference_loop.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
