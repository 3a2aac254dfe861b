"""ctypes binding of libleniax_b200.so (C ABI: include/leniax_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Nothing here falls back to another
implementation: a missing library or a missing GPU raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

LNX_MAX_CHANNELS = 8
LNX_MAX_KERNELS = 32
LNX_NB_STATS = 11
LNX_MAX_RINGS = 8
LNX_COUT_ANY, LNX_COUT_NONE = -1, -2

LNX_RUN_EARLY_STOP = 1
LNX_RUN_ASSUME_FINITE = 0x100
LNX_RUN_GENERIC_OLD = 0x400
LNX_RUN_TILED_GENERIC = 0x800
LNX_RUN_GENERIC_1CTA = 0x1000
LNX_RUN_WEIGHTS_MATCH_COUT = 0x2000
LNX_RUN_T64_LINE = 0x4000
LNX_RUN_T2K_REAL_ROWS = 0x8000
LNX_RUN_T64_STEPWISE = 0x10000
LNX_RUN_T64_WHOLE_SCAN = 0x20000
LNX_PLAN_FORCE_TILED = 1

LNX_OK, LNX_ERR_INVALID, LNX_ERR_UNSUPPORTED, LNX_ERR_CUDA, LNX_ERR_NO_DEVICE = 0, -1, -2, -3, -4

# order of the scalar statistics planes written by lnx_run_scan (lnx_stat_key)
STAT_KEYS = ('mass', 'mass_volume', 'mass_density', 'growth', 'growth_volume', 'growth_density', 'mass_speed',
             'mass_angle_speed', 'mass_growth_dist', 'inertia', 'potential_volume')


class LeniaxB200Error(RuntimeError):
    pass


class LnxDesc(Structure):
    _fields_ = [
        ('nb_dims', c_int32),
        ('dims', c_int32 * 3),
        ('nb_channels', c_int32),
        ('nb_kernels', c_int32),
        ('nb_slots', c_int32),
        ('slot', c_int32 * LNX_MAX_KERNELS),
        ('c_in', c_int32 * LNX_MAX_KERNELS),
        ('gf_id', c_int32 * LNX_MAX_KERNELS),
        ('c_out', c_int32 * LNX_MAX_KERNELS),
        ('state_fn', c_int32),
        ('weighted_average', c_int32),
        ('R', c_float),
        ('stats_dt', c_float),
        ('flags', c_uint32),
    ]


class LnxKernelSpec(Structure):
    _fields_ = [
        ('shape', c_int32),
        ('kf', c_int32),
        ('nb_b', c_int32),
        ('r', c_float),
        ('bs', c_float * LNX_MAX_RINGS),
        ('kf_params', c_float * 2),
        ('a', c_float),
        ('b', c_float),
        ('cos_theta', c_float),
        ('sin_theta', c_float),
    ]


KSHAPE_IDS = {'circle_2d': 1, 'ellipse_2d': 2, 'oriented_ellipse_2d': 3}
KF_IDS = {'poly_quad': 0, 'gauss_bump': 1, 'step': 2, 'gauss': 3, 'threshold': 4, 'staircase': 5, 'triangle': 6}

EXPORTS = {
    'lnx_version': (ctypes.c_int, []),
    'lnx_last_error': (c_char_p, []),
    'lnx_device_count': (ctypes.c_int, []),
    'lnx_plan_create': (ctypes.c_int, [POINTER(LnxDesc), POINTER(c_void_p)]),
    'lnx_plan_destroy': (ctypes.c_int, [c_void_p]),
    'lnx_kernel_table_bytes': (c_size_t, [c_void_p]),
    'lnx_kernels_prepare': (ctypes.c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    'lnx_rfft2': (ctypes.c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    'lnx_measure_fp32_peak': (ctypes.c_int, [c_int32, POINTER(ctypes.c_double), POINTER(ctypes.c_double), c_void_p]),
    'lnx_workspace_bytes': (c_size_t, [c_void_p]),
    'lnx_workspace_bytes_for': (c_size_t, [c_void_p, c_int32, c_int32]),
    'lnx_rfftn': (ctypes.c_int, [c_int32, POINTER(c_int32), c_int32, c_void_p, c_void_p, c_void_p]),
    'lnx_run_scan': (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_uint32] + [c_void_p] * 12 + [c_void_p, c_size_t, c_void_p]),
    'lnx_compute_stats': (ctypes.c_int, [c_void_p, c_int32] + [c_void_p] * 8 + [c_void_p]),
    'lnx_run_scan_variant': (c_char_p, [c_void_p, c_int32]),
    'lnx_summarize_stats': (ctypes.c_int, [POINTER(c_void_p), c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    'lnx_update': (ctypes.c_int, [c_void_p, c_int32] + [c_void_p] * 9),
    'lnx_rasterize_kernels': (ctypes.c_int, [c_int32, POINTER(LnxKernelSpec), c_float, c_int32, c_void_p, c_void_p]),
    'lnx_kernel_spectrum': (ctypes.c_int, [c_int32, POINTER(c_int32), c_int32, POINTER(c_int32), c_void_p, c_void_p, c_void_p]),
    'lnx_random_uniform': (ctypes.c_int, [c_uint64, c_int64, c_void_p, c_void_p]),
    'lnx_init_uniform': (ctypes.c_int, [c_uint64, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    'lnx_init_perlin': (ctypes.c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'lnx_init_perlin_seeded': (ctypes.c_int, [c_int32, POINTER(c_uint64), c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    'lnx_gen2_schedule': (ctypes.c_int, [POINTER(LnxDesc)] + [POINTER(c_int32)] * 4),
    'lnx_update_conv': (ctypes.c_int, [POINTER(LnxDesc), c_int32, c_int32, c_int32] + [c_void_p] * 4 + [c_float] + [c_void_p] * 4),
}

_LIB = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libleniax_b200.so')


def load_library():
    """Load libleniax_b200.so (once).  Raises LeniaxB200Error when it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise LeniaxB200Error(
                f'{path} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). leniax_b200 has no CPU fallback.'
            )
        lib = ctypes.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(status: int):
    """Map an lnx_status to the exception the reference would raise for the same misuse."""
    if status == LNX_OK:
        return
    msg = load_library().lnx_last_error().decode('utf-8', 'replace')
    if status == LNX_ERR_INVALID:
        raise ValueError(msg)
    if status == LNX_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise LeniaxB200Error(msg)
