"""Build libdesilike_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, 'csrc')
LIBDIR = os.path.join(ROOT, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libdesilike_b200.so')
SOURCES = ['plan.cu', 'contract.cu', 'contract_masked.cu', 'features.cu', 'theory_kernels.cu', 'solve.cu', 'peaks.cu', 'sampler.cu', 'mlp.cu', 'gather.cu', 'ozaki.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
              '-cudart', 'static']


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIBPATH):
        return True
    mtime = os.path.getmtime(LIBPATH)
    deps = sources() + [os.path.join(CSRC, 'common.h'), os.path.join(ROOT, '..', 'include', 'desilike_b200.h')]
    return any(os.path.getmtime(d) > mtime for d in deps if os.path.exists(d))


def _compile(nvcc, flags, objdir, out, force, verbose, extra=()):
    """Per-source objects compiled in parallel (only the stale ones), then one link step."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, 'common.h'), os.path.join(CSRC, 'contract.cu'), os.path.join(ROOT, '..', 'include', 'desilike_b200.h')]
    hm = max(os.path.getmtime(h) for h in headers if os.path.exists(h))
    cflags = [f for f in flags if f not in ('-shared',)]
    jobs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            jobs.append((src, obj))

    def one(job):
        cmd = [nvcc] + cflags + list(extra) + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', job[1], job[0]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed:\n{}\n{}'.format(' '.join(cmd), res.stderr))
        return res.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(len(jobs), 1))) as pool:
        logs = list(pool.map(one, jobs))
    if verbose:
        print('\n'.join(logs))
    objs = [os.path.join(objdir, os.path.basename(src)[:-3] + '.o') for src in sources()]
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-cudart', 'static', '-o', out] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc link failed:\n{}\n{}'.format(' '.join(cmd), res.stderr))
    return out


def build(force=False, verbose=False):
    """Compile every CUDA source into ``desilike_b200/lib/libdesilike_b200.so``.  nvcc cross-compiles without a GPU."""
    if not force and not needs_build():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    return _compile(nvcc, NVCC_FLAGS, os.path.join(LIBDIR, 'obj'), LIBPATH, force, verbose)


def build_profile():
    """Phase-profiling build of the same sources (-DDLB_PROFILE: per-phase cycle counters in the contraction kernel) ->
    lib/libdesilike_b200_prof.so, loaded by scripts/phase_profile.py through DESILIKE_B200_LIB."""
    out = os.path.join(LIBDIR, 'libdesilike_b200_prof.so')
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    return _compile(nvcc, NVCC_FLAGS, os.path.join(LIBDIR, 'obj_prof'), out, False, False, extra=['-DDLB_PROFILE'])


if __name__ == '__main__':
    import sys
    if '--profile' in sys.argv:
        print(build_profile())
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
