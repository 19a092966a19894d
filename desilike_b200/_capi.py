"""ctypes binding of include/desilike_b200.h -- the only way the Python host reaches the CUDA path.

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, calls raise.
"""
import ctypes
import os

import numpy as np

from .build import LIBPATH

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

DLB_ABI_VERSION = 10
DLB_COEF_MAXF = 4
GEN_TAYLOR, GEN_MLP, GEN_KAISER, GEN_BAO = 1, 2, 3, 4
PT_NONFINITE, PT_OUT_OF_PRIOR, PT_NOT_SPD = 1, 2, 4


class GeneratorDesc(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('nparams', ctypes.c_int32), ('param_idx', c_int32_p), ('param_default', c_double_p),
                ('nfeat', ctypes.c_int32), ('ni', ctypes.c_int32), ('idata', c_int32_p), ('nd', ctypes.c_int64), ('ddata', c_double_p)]


class CoefDesc(ctypes.Structure):
    _fields_ = [('scale', ctypes.c_double), ('nf', ctypes.c_int32), ('idx', ctypes.c_int32 * DLB_COEF_MAXF), ('pow', ctypes.c_int32 * DLB_COEF_MAXF)]


class TermDesc(ctypes.Structure):
    _fields_ = [('slot', ctypes.c_int32), ('coef', ctypes.c_int32), ('gen', ctypes.c_int32), ('foff', ctypes.c_int32), ('K', ctypes.c_int32),
                ('moff', ctypes.c_int64), ('moff_plain', ctypes.c_int64)]


class BlockDesc(ctypes.Structure):
    _fields_ = [('ncols', ctypes.c_int32), ('nslots', ctypes.c_int32), ('nterms', ctypes.c_int32), ('terms', ctypes.POINTER(TermDesc)),
                ('matrices_white', c_double_p), ('matrices_plain', c_double_p), ('r0_white', c_double_p), ('t0_plain', c_double_p),
                ('slot_solved', c_int32_p), ('nconst', ctypes.c_int32), ('const_solved', c_int32_p), ('gconst_white', c_double_p),
                ('transform_data', c_double_p), ('whitening_t', c_double_p), ('r0_data_white', c_double_p), ('ncols_white', ctypes.c_int32),
                ('entry_kind', c_int32_p), ('entry_i', c_int32_p), ('entry_j', c_int32_p), ('skip_const_gram', ctypes.c_int32),
                ('q_ncols', ctypes.c_int32), ('q_ngroups', ctypes.c_int32), ('q_matrix', c_double_p), ('q_tiles', c_int32_p), ('q_groups', c_int32_p),
                ('q_dent', c_int32_p), ('q_const', c_double_p), ('q_foff', ctypes.c_int32), ('q_K', ctypes.c_int32), ('q_nw', ctypes.c_int32),
                ('q_wconst_val', c_double_p), ('q_wconst_idx', c_int32_p)]


class PlanDesc(ctypes.Structure):
    _fields_ = [('abi_version', ctypes.c_int32), ('d', ctypes.c_int32),
                ('prior_kind', c_int32_p), ('prior_lo', c_double_p), ('prior_hi', c_double_p), ('prior_loc', c_double_p), ('prior_scale', c_double_p),
                ('ngen', ctypes.c_int32), ('gens', ctypes.POINTER(GeneratorDesc)),
                ('ncoef', ctypes.c_int32), ('coefs', ctypes.POINTER(CoefDesc)),
                ('nblocks', ctypes.c_int32), ('blocks', ctypes.POINTER(BlockDesc)),
                ('ns', ctypes.c_int32), ('nmarg', ctypes.c_int32), ('x0', c_double_p), ('xprior_kind', c_int32_p),
                ('xprior_loc', c_double_p), ('xprior_scale', c_double_p), ('xprior_lo', c_double_p), ('xprior_hi', c_double_p),
                ('nderived', ctypes.c_int32), ('derived_off', c_int32_p), ('derived_code', c_int32_p), ('derived_const', c_double_p),
                ('m0_extra', c_double_p), ('mlp_slices', ctypes.c_int32)]


# every symbol include/desilike_b200.h declares: (restype, argtypes)
SYMBOLS = {
    'dlb_abi_version': (ctypes.c_int, []),
    'dlb_last_error': (ctypes.c_char_p, []),
    'dlb_device_count': (ctypes.c_int, []),
    'dlb_plan_create': (ctypes.c_int, [ctypes.POINTER(PlanDesc), ctypes.POINTER(ctypes.c_void_p)]),
    'dlb_plan_destroy': (None, [ctypes.c_void_p]),
    'dlb_plan_ntheory': (ctypes.c_int, [ctypes.c_void_p]),
    'dlb_plan_nsolved': (ctypes.c_int, [ctypes.c_void_p]),
    'dlb_plan_flags': (ctypes.c_int, [ctypes.c_void_p]),
    'dlb_plan_executed_flops': (ctypes.c_double, [ctypes.c_void_p]),
    'dlb_eval_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 5 + [ctypes.c_void_p]),
    'dlb_eval_batch_hessian': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 6 + [ctypes.c_void_p]),
    'dlb_theory_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    'dlb_features_batch': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    'dlb_ensemble_step': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    'dlb_launch_count': (ctypes.c_int64, []),
    'dlb_measure_fp64_peak': (ctypes.c_int, [ctypes.c_int32, c_double_p]),
    'dlb_last_kernel_ms': (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]),
    'dlb_set_timing': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    'dlb_last_fftlog_ms': (ctypes.c_int, [c_double_p]),
    'dlb_ozaki_matmul': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_void_p]),
    'dlb_gather_create': (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    'dlb_gather_ipc_handle': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'dlb_gather_connect': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'dlb_gather_buffer': (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int32]),
    'dlb_eval_batch_gather': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'dlb_gather_publish': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'dlb_gather_wait': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    'dlb_gather_status': (ctypes.c_int, [ctypes.c_void_p]),
    'dlb_gather_set_timeout': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double]),
    'dlb_gather_destroy': (None, [ctypes.c_void_p]),
}

_lib = None


class DesilikeB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (built by ``desilike_b200.build.build()``); raise if it is missing -- no fallback."""
    global _lib
    if _lib is None:
        libpath = os.environ.get('DESILIKE_B200_LIB', LIBPATH)   # override: e.g. a -DDLB_PROFILE build of the same sources
        if not os.path.exists(libpath):
            raise DesilikeB200Error('{} not found: run `python -m desilike_b200.build` (or __graft_entry__.build()); '
                                    'desilike_b200 has no CPU fallback'.format(libpath))
        lib = ctypes.CDLL(libpath)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        if lib.dlb_abi_version() != DLB_ABI_VERSION:
            raise DesilikeB200Error('ABI version mismatch between {} and the Python host'.format(LIBPATH))
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise DesilikeB200Error('libdesilike_b200 error {:d}: {}'.format(rc, load().dlb_last_error().decode()))


def as_ptr(array, ctype=ctypes.c_double):
    if array is None:
        return None
    return array.ctypes.data_as(ctypes.POINTER(ctype))


def f64(array):
    return np.ascontiguousarray(array, dtype=np.float64)


def i32(array):
    return np.ascontiguousarray(array, dtype=np.int32)
