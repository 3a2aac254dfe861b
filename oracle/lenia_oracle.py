"""NumPy restatement of the leniax simulation hot path (TEST INFRASTRUCTURE).

This module is the *oracle*: a plain NumPy/scipy.fft restatement of what the
reference (morgangiraud/leniax, JAX) computes on the path named by
BASELINE.json.  It is used only as a checker by ``tests/``, by
``__graft_entry__.smoke()`` and as the CPU baseline of ``bench.py``.  It is never
imported by the product package ``leniax_b200``.

Parity status: **pinned** on the state trajectory — it reproduces the three
golden last-frame fixtures of the reference's own test-suite
(``tests/test_pipeline.py:18-130`` with ``tests/fixtures/*_last_frame*.p``; see
``tests/test_oracle_golden.py``) and the known-answer tests of
``tests/test_core.py:55-103,159-185``, ``tests/test_statistics.py:14-50`` and
``tests/test_kernels.py:12-20``.  **Unpinned** (no reference fixture exists):
the numeric values of the 12 statistics and of ``stats['N']`` on real runs, and
the perlin / uniform initial states (``perlin.py``, ``initializations.py``:
restated from the angles / uniform draw onwards, no reference test fixes one).

Third-party arithmetic that is not in /root/reference: the reference lowers
``jnp.fft.fftn`` / reductions through JAX/XLA (pinned env jax 0.2.26 / jaxlib
0.1.75, ``environment_linux.yml:57-58``); on CPU that is pocketfft in complex64.
Here ``scipy.fft`` (also pocketfft, keeps complex64) plays that role.  Also
third-party: ``scipy.ndimage.zoom`` (``helpers.py:61``, called directly here, as
the reference does) and pyribs ``ribs==0.4.0`` ``GridArchive.get_index``
(``setup.py:21``; restated in ``grid_archive_index`` from its published source,
no reference test pins an archive index: unpinned).

Every function carries the reference file:line it follows.  ``dtype`` selects
the float32 oracle (default, what JAX computes) or its float64 twin (used to
measure the fp32 noise floor).
"""
from __future__ import annotations

import base64
import copy
import gzip
import math
from fractions import Fraction
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft as sfft

# leniax/constant.py:6-13
EPSILON = 1e-7
START_CHECK_STOP = 10
NB_STATS_STEPS = 128
NB_CHARS = (ord('Z') - ord('A')) + (ord('z') - ord('a')) + (ord('þ') - ord('À'))  # 112

# leniax/statistics.py:284,313-314
MONOTONIC_STOP_STEP = 128
MASS_VOLUME_THRESHOLD = 10.
MASS_VOLUME_STOP_STEP = 128

STAT_KEYS = (
    'mass', 'mass_volume', 'mass_density', 'growth', 'growth_volume', 'growth_density', 'mass_speed',
    'mass_angle_speed', 'mass_growth_dist', 'inertia', 'potential_volume'
)  # + 'channel_mass' [.., C]  (leniax/statistics.py:102-115)


def _cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


# ---------------------------------------------------------------------------
# kernel shell functions                           leniax/kernel_functions.py
# ---------------------------------------------------------------------------
def kf_poly_quad(params, X):  # kernel_functions.py:7-37
    q = params[0]
    return (4 * X * (1 - X))**q


def kf_gauss_bump(params, X):  # kernel_functions.py:40-69
    q = params[0]
    return np.exp(q * (q - 1 / (X * (1 - X) + EPSILON)))


def kf_step(params, X):  # kernel_functions.py:72-102
    q = params[0]
    return (X >= q) * (X <= 1 - q)


def kf_gauss(params, X):  # kernel_functions.py:105-134
    q = params[0]
    return np.exp(-(((X - q) / (0.3 * q))**2) / 2)


def kf_threshold(params, X):  # kernel_functions.py:137-166
    return X >= params[0]


def _staircase(params, X):  # kernel_functions.py:169-205 / growth_functions.py:163-201
    m, s = params[0], params[1]
    out = 0.5 * (X >= m - s) * (X < m - s / 2)
    out = out + 1 * (X >= m - s / 2) * (X <= m + s / 2)
    out = out + 0.5 * (X > m + s / 2) * (X <= m + s)
    return out


def _triangle(params, X):  # kernel_functions.py:208-250 / growth_functions.py:204-243
    m, s = params[0], params[1]
    left, right = m - s, m + s
    out = (X >= left) * (X < m) * (X - left) / (m - left)
    out = out + (X >= m) * (X <= right) * (X - right) / (m - right)
    return out


KERNEL_FUNCTIONS: Dict[str, Callable] = {  # kernel_functions.py:253-261
    'poly_quad': kf_poly_quad,
    'gauss_bump': kf_gauss_bump,
    'gauss': kf_gauss,
    'step': kf_step,
    'threshold': kf_threshold,
    'staircase': _staircase,
    'triangle': _triangle,
}


# ---------------------------------------------------------------------------
# growth functions                                 leniax/growth_functions.py
# ---------------------------------------------------------------------------
def gf_poly_quad4(params, X):  # growth_functions.py:6-45
    m, s = params[0], params[1]
    with np.errstate(divide='ignore', invalid='ignore'):
        out = 1 - (X - m)**2 / (9 * s**2)
    out = np.maximum(0, out)  # NaN propagates, as jnp.maximum does
    return 2 * out**4 - 1


def gf_gaussian(params, X):  # growth_functions.py:48-82
    m, s = params[0], params[1]
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.exp(-(((X - m) / s)**2) / 2)
    return 2 * out - 1


def gf_gaussian_target(params, X):  # growth_functions.py:85-119
    m, s = params[0], params[1]
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.exp(-(((X - m) / s)**2) / 2)


def gf_step(params, X):  # growth_functions.py:122-160
    m, s = params[0], params[1]
    return 2 * (np.abs(X - m) <= s) - 1


def gf_staircase(params, X):  # growth_functions.py:163-201
    return 2 * _staircase(params, X) - 1


def gf_triangle(params, X):  # growth_functions.py:204-243
    with np.errstate(divide='ignore', invalid='ignore'):
        return 2 * _triangle(params, X) - 1


def gf_identity(params, X):  # growth_functions.py:246-253
    return X


GROWTH_FUNCTIONS: Dict[str, Callable] = {  # growth_functions.py:256-264
    'poly_quad4': gf_poly_quad4,
    'gaussian': gf_gaussian,
    'gaussian_target': gf_gaussian_target,
    'step': gf_step,
    'staircase': gf_staircase,
    'triangle': gf_triangle,
    'identity': gf_identity,
}


# ---------------------------------------------------------------------------
# kernel rasterisation + packing                          leniax/kernels.py
# ---------------------------------------------------------------------------
class KernelMapping:
    """leniax/kernels.py:10-63 (same attribute names)."""
    def __init__(self, nb_channels: int, nb_kernels: int):
        self.cin_kernels: List[List[int]] = [[] for _ in range(nb_channels)]
        self.cin_k_params: List[List] = [[] for _ in range(nb_channels)]
        self.cin_kfs: List[List[str]] = [[] for _ in range(nb_channels)]
        self.cin_gfs: List[List[str]] = [[] for _ in range(nb_channels)]
        self.cin_gf_params: List[List] = [[] for _ in range(nb_channels)]
        self.kernels_weight_per_channel = [[0.] * nb_kernels for _ in range(nb_channels)]
        self.true_channels: Optional[List[bool]] = None

    def get_gf_params(self, dtype=np.float32) -> np.ndarray:
        return np.array([p for sub in self.cin_gf_params for p in sub], dtype=dtype)

    def get_kernels_weight_per_channel(self, dtype=np.float32) -> np.ndarray:
        return np.array(self.kernels_weight_per_channel, dtype=dtype)


def _radial_profile(distances, bs, kf_slug, kf_params, dtype):
    """Shared tail of circle_2d / ellipse_2d: kernels.py:197-207."""
    nb_b = bs.shape[0]
    B_dist = (nb_b * distances).astype(dtype)
    ring = bs[np.minimum(np.floor(B_dist).astype(np.int64), nb_b - 1)]
    shell = KERNEL_FUNCTIONS[kf_slug](kf_params, B_dist % 1)
    return ((distances < 1) * shell * ring).astype(dtype)


def _centered_grid(k_radius_px: int, scale: float, dtype):
    coords = np.indices([2 * k_radius_px, 2 * k_radius_px]) - k_radius_px  # kernels.py:190-193
    return (coords / scale).astype(dtype)


def circle_2d(R, k_params, kf_slug, kf_params, dtype=np.float32):  # kernels.py:176-212
    r = k_params[0]
    bs = np.array(k_params[1], dtype=dtype)
    k_radius_px = math.ceil(r * R)
    cc = _centered_grid(k_radius_px, r * R, dtype)
    distances = np.sqrt(np.sum(cc**2, axis=0)).astype(dtype)
    kernel = _radial_profile(distances, bs, kf_slug, kf_params, dtype)
    kernel = kernel / kernel.sum(dtype=dtype)
    return kernel[np.newaxis].astype(dtype)


def _rotated(cc, theta, dtype):  # kernels.py:236-239
    c, s = dtype(np.cos(theta)), dtype(np.sin(theta))
    return np.stack([cc[0] * c + cc[1] * s, -cc[0] * s + cc[1] * c]).astype(dtype)


def ellipse_2d(R, k_params, kf_slug, kf_params, dtype=np.float32):  # kernels.py:215-262
    r = k_params[0]
    k_radius_px = math.ceil(r * R)
    bs = np.array(k_params[1], dtype=dtype)
    a, b, theta = k_params[2], k_params[3], k_params[4] * np.pi
    rc = _rotated(_centered_grid(k_radius_px, r * R, dtype), theta, dtype)
    distances = np.sqrt((rc[0] / a)**2 + (rc[1] / b)**2).astype(dtype)
    kernel = _radial_profile(distances, bs, kf_slug, kf_params, dtype)
    kernel = kernel / kernel.sum(dtype=dtype)
    grad = rc[0].copy()
    grad[rc[0] < -0.01] = -1  # kernels.py:256-257 (second mask is evaluated on the updated array)
    grad[grad > 0.01] = 1
    return (kernel * grad)[np.newaxis].astype(dtype)


def oriented_ellipse_2d(R, k_params, kf_slug, kf_params, dtype=np.float32):  # kernels.py:265-309
    r = k_params[0]
    k_radius_px = math.ceil(r * R)
    bs = np.array(k_params[1], dtype=dtype)
    a, b, theta = k_params[2], k_params[3], k_params[4] * np.pi
    rc = _rotated(_centered_grid(k_radius_px, r * R, dtype), theta, dtype)
    distances = np.sqrt((rc[0] / a)**2 + (rc[1] / b)**2).astype(dtype)
    kernel = _radial_profile(distances, bs, kf_slug, kf_params, dtype) * rc[0]
    kernel = kernel / np.abs(kernel).sum(dtype=dtype)
    return kernel[np.newaxis].astype(dtype)


def raw(R, k_params, kf_slug, kf_params, dtype=np.float32):  # kernels.py:161-173
    return np.array(k_params, dtype=dtype)


KERNEL_SHAPES: Dict[str, Callable] = {  # kernels.py:312-317
    'raw': raw,
    'circle_2d': circle_2d,
    'ellipse_2d': ellipse_2d,
    'oriented_ellipse_2d': oriented_ellipse_2d,
}


def crop_zero(kernels: np.ndarray) -> np.ndarray:  # utils.py:296-318
    nz = kernels != 0
    if kernels.ndim == 3:
        keep1 = nz.any(axis=(0, 2))
        keep2 = nz.any(axis=(0, 1))
        return kernels[:, keep1][:, :, keep2]
    if kernels.ndim == 4:
        keep1 = nz.any(axis=(0, 2, 3))
        keep2 = nz.any(axis=(0, 1, 3))
        keep3 = nz.any(axis=(0, 1, 2))
        return kernels[:, keep1][:, :, keep2][:, :, :, keep3]
    raise ValueError("Can't handle more than 3 dimensions")


def get_kernels_and_mapping(kernels_params: List, world_size: Sequence[int], nb_channels: int, R: float,
                            fft: bool = True, dtype=np.float32) -> Tuple[np.ndarray, KernelMapping]:
    """leniax/kernels.py:66-158.  Sorts ``kernels_params`` in place, like the reference."""
    world_size = list(world_size)
    mapping = KernelMapping(nb_channels, len(kernels_params))
    kernels_params.sort(key=lambda d: d['c_in'])  # kernels.py:90 (stable, in place)
    padded = []
    for idx, p in enumerate(kernels_params):
        k = KERNEL_SHAPES[p['k_slug']](R, p['k_params'], p['kf_slug'], p['kf_params'], dtype)
        pads = [(0, 0)]
        for ws, ks in zip(world_size, k.shape[1:]):  # kernels.py:93-100
            lo = (ws - ks) // 2
            pads.append((lo, lo if (ws - ks) % 2 == 0 else lo + 1))
        padded.append(np.pad(k, pads))
        mapping.cin_kernels[p['c_in']].append(idx)
        mapping.cin_gfs[p['c_in']].append(p['gf_slug'])
        mapping.cin_gf_params[p['c_in']].append(p['gf_params'])
        mapping.cin_kfs[p['c_in']].append(p['kf_slug'])
        mapping.cin_k_params[p['c_in']].append(p['k_params'])
        mapping.kernels_weight_per_channel[p['c_out']][idx] = p['h']

    kernels = np.vstack(padded)  # [nb_kernels, *dims]
    if not fft:
        kernels = crop_zero(kernels)
    kshape = kernels.shape[1:]

    max_k = max(len(l) for l in mapping.cin_kernels)
    per_channel, true_channels = [], []
    for lst in mapping.cin_kernels:  # kernels.py:122-143
        kc = kernels[np.array(lst, dtype=np.int64)] if len(lst) else np.zeros((0, ) + kshape, dtype=dtype)
        missing = max_k - kc.shape[0]
        true_channels += [True] * kc.shape[0] + [False] * missing
        if missing:
            kc = np.concatenate([kc, np.zeros((missing, ) + kshape, dtype=dtype)])
        per_channel.append(kc)
    mapping.true_channels = None if all(true_channels) else true_channels

    if fft:
        axes = tuple(range(-len(world_size), 0))
        K = np.stack(per_channel)[np.newaxis]  # [1, C, max_k, *dims]
        K = sfft.fftn(sfft.fftshift(K, axes=axes), axes=axes).astype(_cdtype(dtype))
    else:
        K = np.concatenate(per_channel)[:, np.newaxis]  # [C*max_k, 1, kh, kw]
    return K, mapping


def tc_indices_of(mapping: KernelMapping) -> Optional[Tuple[int, ...]]:  # helpers.py:449-456
    if mapping.true_channels is None:
        return None
    return tuple(i for i, t in enumerate(mapping.true_channels) if t)


# ---------------------------------------------------------------------------
# one step                                                   leniax/core.py
# ---------------------------------------------------------------------------
def get_potential_fft(state: np.ndarray, K: np.ndarray, tc_indices=None) -> np.ndarray:
    """core.py:52-102 (channel_first).  state [N,C,*dims], K [1,C,max_k,*dims] complex."""
    nd = state.ndim - 2
    axes = tuple(range(-nd, 0))
    spec = sfft.fftn(state, axes=axes)[:, :, np.newaxis]  # [N, C, 1, *dims]
    conv = np.real(sfft.ifftn(spec * K, axes=axes))  # [N, C, max_k, *dims]
    conv = conv.reshape((-1, K.shape[2] * state.shape[1]) + state.shape[2:])
    if tc_indices is not None:
        conv = np.take(conv, np.array(tc_indices), axis=1)
    return conv.astype(state.dtype)


def get_potential_conv(state: np.ndarray, K: np.ndarray, tc_indices=None) -> np.ndarray:
    """core.py:105-146 + helpers.py:464-488: wrap-pad then VALID depthwise cross-correlation (2-D).

    state [N,C,H,W]; K [C*max_k, 1, kh, kw] (real).  Used as an FFT-independent cross-check.
    """
    N, C, H, W = state.shape
    O, _, kh, kw = K.shape
    per = O // C
    pad = []
    for d in (kh, kw):
        pad.append((d // 2, d // 2 - 1) if d % 2 == 0 else (d // 2, d // 2))
    padded = np.pad(state, [(0, 0), (0, 0)] + pad, mode='wrap')
    out = np.zeros((N, O, H, W), dtype=state.dtype)
    for o in range(O):
        c = o // per
        acc = np.zeros((N, H, W), dtype=state.dtype)
        for i in range(kh):
            for j in range(kw):
                w = K[o, 0, i, j]
                if w != 0:
                    acc += w * padded[:, c, i:i + H, j:j + W]
        out[:, o] = acc
    if tc_indices is not None:
        out = np.take(out, np.array(tc_indices), axis=1)
    return out


def weighted_sum(fields: np.ndarray, weights: np.ndarray) -> np.ndarray:  # core.py:202-225
    return np.einsum('ck,nk...->nc...', weights, fields).astype(fields.dtype)


def weighted_mean(fields: np.ndarray, weights: np.ndarray) -> np.ndarray:  # core.py:228-242
    out = weighted_sum(fields, weights)
    shape = (1, -1) + (1, ) * (fields.ndim - 2)
    with np.errstate(divide='ignore', invalid='ignore'):
        return out / weights.sum(axis=1).reshape(shape)


def get_field(potential, gf_params, weights, gf_slugs: Sequence[str], average: bool = True):
    """core.py:163-199 + helpers.py:491-515.  ``gf_slugs`` flat in kernel order."""
    subs = [GROWTH_FUNCTIONS[s](gf_params[i], potential[:, i]) for i, s in enumerate(gf_slugs)]
    fields = np.stack(subs, axis=1).astype(potential.dtype)
    return (weighted_mean if average else weighted_sum)(fields, weights).astype(potential.dtype)


def get_state_v1(state, field, dt):  # core.py:245-272 (forward value of the straight-through estimator)
    return np.clip(state + dt * field, 0., 1.).astype(state.dtype)


def get_state_v2(state, field, dt):  # core.py:275-297
    return (state * (1 - dt) + dt * field).astype(state.dtype)


def get_state_simple(state, field, dt):  # core.py:300-319
    return (state + dt * field).astype(state.dtype)


STATE_FUNCTIONS = {'v1': get_state_v1, 'v2': get_state_v2, 'simple': get_state_simple}  # core.py:322-326


def build_update_fn(mapping: KernelMapping, get_state_fn_slug: str = 'v1', average_weight: bool = True,
                    fft: bool = True) -> Callable:
    """helpers.py:401-427 → callable(state, K, gf_params, W, dt) -> (state', field, potential) (core.py:13-49)."""
    tci = tc_indices_of(mapping)
    gf_slugs = [s for sub in mapping.cin_gfs for s in sub]
    state_fn = STATE_FUNCTIONS[get_state_fn_slug]
    pot_fn = get_potential_fft if fft else get_potential_conv

    def update(state, K, gf_params, weights, dt):
        potential = pot_fn(state, K, tci)
        field = get_field(potential, gf_params, weights, gf_slugs, average_weight)
        return state_fn(state, field, dt), field, potential

    return update


# ---------------------------------------------------------------------------
# statistics                                            leniax/statistics.py
# ---------------------------------------------------------------------------
def center_world(x: np.ndarray, shift_idx: np.ndarray) -> np.ndarray:
    """utils.py:269-293: per-world ``roll(x[n], -shift[n], world axes)``."""
    nd = shift_idx.shape[1]
    axes = tuple(range(-nd, 0))
    out = np.empty_like(x)
    for n in range(x.shape[0]):
        out[n] = np.roll(x[n], tuple(-int(s) for s in shift_idx[n]), axes)
    return out


def build_compute_stats_fn(world_params: Dict, render_params: Dict, dtype=np.float32) -> Callable:
    """leniax/statistics.py:11-128."""
    world_size = list(render_params['world_size'])
    nd = len(world_size)
    R = world_params['R']
    dt = 1. / world_params['T']
    R2 = dtype(R**2)
    Rf = dtype(R)
    dtf = dtype(dt)
    eps = dtype(EPSILON)
    world_axes = tuple(range(-nd, 0))
    nb_axes = tuple(range(-(1 + nd), 0))
    mid = np.array([s // 2 for s in world_size]).reshape((nd, ) + (1, ) * nd)
    cc = (np.indices(world_size) - mid).astype(dtype)  # [D, *dims]
    cc = cc.reshape((nd, 1, 1) + tuple(world_size))

    def compute_stats(cells, field, potential, prev_shift, prev_centroid, prev_angle):
        ccells = center_world(cells, prev_shift)
        cfield = center_world(field, prev_shift)
        pos_field = np.maximum(cfield, 0)

        m_00 = ccells.sum(axis=nb_axes, dtype=dtype)
        g_00 = pos_field.sum(axis=nb_axes, dtype=dtype)
        potential_volume = ((potential > eps).sum(axis=nb_axes) / R2).astype(dtype)
        channel_mass = (ccells.sum(axis=world_axes, dtype=dtype) / R2).astype(dtype)
        mass = m_00 / R2
        mass_volume = ((ccells > eps).sum(axis=nb_axes) / R2).astype(dtype)
        mass_density = mass / (mass_volume + eps)
        growth = g_00 / R2
        growth_volume = ((pos_field > eps).sum(axis=nb_axes) / R2).astype(dtype)
        growth_density = growth / (growth_volume + eps)

        AX = ccells[np.newaxis] * cc  # [D, N, C, *dims]
        MX = AX.sum(axis=nb_axes, dtype=dtype)  # [D, N]
        mass_centroid = MX / (m_00 + eps)
        delta = mass_centroid - prev_centroid
        dist_m = np.sqrt((delta**2).sum(axis=0, dtype=dtype))
        mass_speed = dist_m / Rf / dtf
        with np.errstate(invalid='ignore'):
            mass_angle = np.degrees(np.arctan2(delta[1], delta[0])).astype(dtype) * (dist_m / Rf > 0.001)
            mass_angle_speed = ((mass_angle - prev_angle + 540) % 360 - 180) / dtf

        GX = (pos_field[np.newaxis] * cc).sum(axis=nb_axes, dtype=dtype)
        growth_centroid = GX / (g_00 + eps)
        mass_growth_dist = np.sqrt(((growth_centroid - mass_centroid)**2).sum(axis=0, dtype=dtype)) / Rf

        MX2 = (AX * cc).sum(axis=nb_axes, dtype=dtype)
        inertia = ((MX2 - mass_centroid * MX) / (m_00**2 + eps)).sum(axis=0, dtype=dtype)

        stats = {
            'channel_mass': channel_mass,
            'mass': mass,
            'mass_volume': mass_volume,
            'mass_density': mass_density,
            'growth': growth,
            'growth_volume': growth_volume,
            'growth_density': growth_density,
            'mass_speed': mass_speed,
            'mass_angle_speed': mass_angle_speed,
            'mass_growth_dist': mass_growth_dist,
            'inertia': inertia,
            'potential_volume': potential_volume,
        }
        stats = {k: np.asarray(v, dtype=dtype) for k, v in stats.items()}

        with np.errstate(invalid='ignore'):
            trunc = np.nan_to_num(mass_centroid, nan=0.0, posinf=0.0, neginf=0.0).astype(np.int32)  # :117 astype
        world_shape = np.array(cells.shape[2:], dtype=np.int32)
        total_shift = (prev_shift + trunc.T) % world_shape  # :119 (sign of the divisor, like Python)
        mass_centroid = (mass_centroid - trunc).astype(dtype)  # :124
        return stats, total_shift.astype(np.int32), mass_centroid, mass_angle.astype(dtype)

    return compute_stats


def monotonic_heuristic(sign, previous_sign, counter):  # statistics.py:287-306
    counter = counter * (sign == previous_sign) + 1
    return counter <= MONOTONIC_STOP_STEP, counter


def mass_volume_heuristic(mass_volume, counter):  # statistics.py:317-333
    counter = counter * (mass_volume > MASS_VOLUME_THRESHOLD) + 1
    return counter <= MASS_VOLUME_STOP_STEP, counter


def check_heuristics(stats: Dict[str, np.ndarray]) -> np.ndarray:
    """statistics.py:134-205.  stats[k] is [T, N] (channel_mass [T, N, C]).  Returns [T, N] float."""
    mass = stats['mass']
    T, N = mass.shape
    dtype = mass.dtype
    eps = dtype.type(EPSILON)
    should_continue = np.ones(N, dtype=dtype)
    init_cm = stats['channel_mass'][0]
    prev_mass = mass[0]
    prev_sign = np.zeros(N, dtype=dtype)
    mono = np.zeros(N, dtype=np.int32)
    vol = np.zeros(N, dtype=np.int32)
    out = np.empty((T, N), dtype=dtype)
    for t in range(T):
        cm = stats['channel_mass'][t]
        cond = (cm >= eps).all(axis=1) * (cm <= 3 * init_cm).all(axis=1)
        with np.errstate(invalid='ignore'):
            sign = np.sign(mass[t] - prev_mass)
        c, mono = monotonic_heuristic(sign, prev_sign, mono)
        cond = cond * c
        c, vol = mass_volume_heuristic(stats['mass_volume'][t], vol)
        cond = cond * c
        should_continue = should_continue * cond
        prev_mass, prev_sign = mass[t], sign
        out[t] = should_continue
    return out


# ---------------------------------------------------------------------------
# time loops                                               leniax/runner.py
# ---------------------------------------------------------------------------
def _init_carry(cells0, dtype):  # runner.py:271-292
    N, nd = cells0.shape[0], cells0.ndim - 2
    return (np.zeros((N, nd), np.int32), np.zeros((nd, N), dtype), np.zeros((N, ), dtype))


def run_scan(cells0, K, gf_params, weights, T, max_run_iter, update_fn, compute_stats_fn,
             keep_intermediary_data: bool = True):
    """runner.py:119-164 (+ _scan_fn 295-334).  Stats are taken on the PRE-update cells."""
    dtype = cells0.dtype
    dt = dtype.type(1.) / dtype.type(T)
    shift, centroid, angle = _init_carry(cells0, dtype)
    cells = cells0
    all_c, all_f, all_p, all_s = [], [], [], []
    for _ in range(max_run_iter):
        new_cells, field, potential = update_fn(cells, K, gf_params, weights, dt)
        st, shift, centroid, angle = compute_stats_fn(cells, field, potential, shift, centroid, angle)
        if keep_intermediary_data:
            all_c.append(cells)
            all_f.append(field)
            all_p.append(potential)
        all_s.append(st)
        cells = new_cells
    stats = {k: np.stack([s[k] for s in all_s]) for k in all_s[0]}
    stats['N'] = check_heuristics(stats).sum(axis=0)  # runner.py:161-162
    if keep_intermediary_data:
        return np.stack(all_c), np.stack(all_f), np.stack(all_p), stats
    return stats, cells


def run_scan_mem_optimized(cells0, K, gf_params, weights, T, max_run_iter, update_fn, compute_stats_fn):
    """runner.py:167-215: leading N_sols axis on cells0/K/gf_params/weights/T.

    Returns (stats {k: [N_sols, T, N_init]}, final_cells [N_sols, N_init, C, *dims]).
    """
    per_sol = [
        run_scan(cells0[i], K[i], gf_params[i], weights[i], T[i], max_run_iter, update_fn, compute_stats_fn, False)
        for i in range(cells0.shape[0])
    ]
    stats = {k: np.stack([s[0][k] for s in per_sol]) for k in per_sol[0][0]}
    return stats, np.stack([s[1] for s in per_sol])


def run(cells, K, gf_params, weights, T, max_run_iter, update_fn, compute_stats_fn, stat_trunc: bool = False):
    """runner.py:16-116: python loop with on-the-fly heuristics (total-mass rules, grace period)."""
    assert max_run_iter > 0 and cells.shape[0] == 1
    dtype = cells.dtype
    dt = dtype.type(1.) / dtype.type(T)
    all_c, all_f, all_p, all_s = [cells], [], [], []
    init_mass = cells.sum(dtype=dtype)
    prev_mass = init_mass
    prev_sign = np.zeros(1, np.int32)
    mono = np.zeros(1, np.int32)
    vol = np.zeros(1, np.int32)
    should_continue = np.ones(1, np.int32)
    shift, centroid, angle = _init_carry(cells, dtype)
    current_iter = 0
    for current_iter in range(max_run_iter):
        new_cells, field, potential = update_fn(cells, K, gf_params, weights, dt)
        st, shift, centroid, angle = compute_stats_fn(cells, field, potential, shift, centroid, angle)
        cells = new_cells
        all_c.append(cells)
        all_f.append(field)
        all_p.append(potential)
        all_s.append(st)
        mass = st['mass']
        cond = (mass >= EPSILON) * (mass <= 3 * init_mass)  # runner.py:84-88
        sign = np.sign(mass - prev_mass)
        c, mono = monotonic_heuristic(sign, prev_sign, mono)
        cond = cond * c
        c, vol = mass_volume_heuristic(st['mass_volume'], vol)
        cond = cond * c
        should_continue = should_continue * cond
        # NB: the reference never updates previous_mass / previous_sign in this loop (runner.py:90-93)
        if stat_trunc and current_iter >= START_CHECK_STOP and int(should_continue[0]) == 0:
            break
    all_c.pop()
    stats = {k: np.stack([s[k] for s in all_s]) for k in all_s[0]}
    stats['N'] = np.array(current_iter)
    return np.stack(all_c), np.stack(all_f), np.stack(all_p), stats


# ---------------------------------------------------------------------------
# QD consumer                                            leniax/qd.py:150-188
# ---------------------------------------------------------------------------
def grid_archive_index(features: np.ndarray, grid_shape: Sequence[int], features_domain: Sequence[Sequence[float]]) -> np.ndarray:
    """ribs 0.4.0 ``GridArchive.get_index`` (third-party, absent from /root/reference; restated from its published source,
    unpinned): clip ``bc + 1e-6`` to ``[lower, upper - 1e-6]`` then ``int((bc - lower) / interval * dims)`` per axis."""
    f = np.asarray(features, np.float64)
    lower = np.array([d[0] for d in features_domain], np.float64)
    upper = np.array([d[1] for d in features_domain], np.float64)
    eps = 1e-6
    f = np.clip(f + eps, lower, upper - eps)
    return ((f - lower) / (upper - lower) * np.array(grid_shape, np.float64)).astype(np.int64)


def behaviours_of(stats: Dict[str, np.ndarray], fitness_coef: float = 1.):
    """qd.py:168-186: fitness = coef*max_init N; behaviours = mean of the last 128 rows before ns."""
    Ns = stats['N']
    best = np.argmax(Ns, axis=1)
    fitness = fitness_coef * Ns.max(axis=1)
    behaviours = []
    for i in range(Ns.shape[0]):
        ns = max(int(fitness[i]), 128)
        behaviours.append({k: stats[k][i, ns - 128:ns, best[i]].mean(axis=0) for k in stats if k != 'N'})
    return fitness, best, behaviours


# ---------------------------------------------------------------------------
# initial states            leniax/perlin.py, leniax/initializations.py, leniax/loader.py:16-30
# ---------------------------------------------------------------------------
def make_array_compressible(cells: np.ndarray) -> np.ndarray:  # loader.py:16-30
    max_val = NB_CHARS**2 - 1
    cells_int32 = np.round(np.asarray(cells, np.float32) * np.float32(max_val)).astype(np.int32)  # jnp.round = half to even, like np.round
    return (cells_int32 / max_val).astype(np.float32)


def perlin_interpolant(t):  # perlin.py:12-13
    return t * t * t * (t * (t * 6 - 15) + 10)


def generate_perlin_noise_2d(angles: np.ndarray, shape: Sequence[int], res: Sequence[int], nb_noise: int = 1) -> np.ndarray:
    """perlin.py:16-71, statement by statement (including the ``diff``-based corner slicing of lines 51-55: the repeated
    gradient image is ``d`` cells larger than ``shape`` on each axis and the four corner gradients are the four ways of
    cropping it).  ``angles [nb_noise, res0, res1]`` float32 -> ``[nb_noise, *shape]`` float32."""
    angles = np.asarray(angles, np.float32)
    gradients = np.stack([np.cos(angles), np.sin(angles)], axis=-1).astype(np.float32)  # :46
    gradients = np.pad(gradients, [(0, 0), (0, 1), (0, 1), (0, 0)], mode='wrap')  # :47
    d = (shape[0] // res[0], shape[1] // res[1])  # :48
    gradients = gradients.repeat(d[0], 1).repeat(d[1], 2)  # :49
    diff = [gradients.shape[1] - shape[0], gradients.shape[2] - shape[1]]  # :51
    g00 = gradients[:, :-diff[0], :-diff[1]]  # :52-55
    g10 = gradients[:, diff[0]:, :-diff[1]]
    g01 = gradients[:, :-diff[0], diff[1]:]
    g11 = gradients[:, diff[0]:, diff[1]:]
    delta = (res[0] / shape[0], res[1] / shape[1])  # :58
    # jnp.mgrid[0:res0:delta0, 0:res1:delta1] (:59) = start + arange(n) * step with n = ceil(res / delta), computed in
    # float32 under jax's default x64-off mode
    n0, n1 = int(math.ceil(res[0] / delta[0])), int(math.ceil(res[1] / delta[1]))
    g0 = (np.arange(n0, dtype=np.float32) * np.float32(delta[0])) % np.float32(1)
    g1 = (np.arange(n1, dtype=np.float32) * np.float32(delta[1])) % np.float32(1)
    grid = np.stack(np.meshgrid(g0, g1, indexing='ij'), axis=-1)[np.newaxis].repeat(nb_noise, 0)  # :59-61
    one = np.float32(1)
    n00 = np.sum(np.stack([grid[..., 0], grid[..., 1]], axis=-1) * g00, 3, dtype=np.float32)  # :62-65
    n10 = np.sum(np.stack([grid[..., 0] - one, grid[..., 1]], axis=-1) * g10, 3, dtype=np.float32)
    n01 = np.sum(np.stack([grid[..., 0], grid[..., 1] - one], axis=-1) * g01, 3, dtype=np.float32)
    n11 = np.sum(np.stack([grid[..., 0] - one, grid[..., 1] - one], axis=-1) * g11, 3, dtype=np.float32)
    t = perlin_interpolant(grid)  # :68
    n0_ = n00 * (one - t[..., 0]) + t[..., 0] * n10  # :69-70
    n1_ = n01 * (one - t[..., 0]) + t[..., 0] * n11
    return (np.float32(np.sqrt(2)) * ((one - t[..., 1]) * n0_ + t[..., 1] * n1_)).astype(np.float32)  # :72


def perlin_from_angles(angles: np.ndarray, world_size: Sequence[int], R: float, gf_params: Sequence[float]) -> np.ndarray:
    """initializations.py:35-77 after the random draw (:63-65): ``angles [nb_init, res0, res1]`` in [0, 2 pi) ->
    ``[nb_init, 1, H, W]`` initial states.  (``jax.random`` itself is third-party and unpinned: SURVEY §8c.)"""
    nb_init = angles.shape[0]
    kernel_radius = math.ceil(R)
    res = [world_size[0] // (kernel_radius * 3), world_size[1] // (kernel_radius * 2)]  # :56
    assert list(angles.shape[1:]) == res, (angles.shape, res)
    lo = gf_params[0]
    hi = min(1, 3 * lo)  # :57-58
    scaling = np.array([lo + i / nb_init * (hi - lo) for i in range(nb_init)], dtype=np.float32)[:, None, None]  # :59-61
    cells = generate_perlin_noise_2d(angles, tuple(world_size), tuple(res), nb_init)
    cells = cells - cells.min(axis=(1, 2), keepdims=True)  # :70-72
    cells = cells / cells.max(axis=(1, 2), keepdims=True)
    cells = cells * scaling
    return make_array_compressible(cells[:, np.newaxis])  # :73-75


def cropped_perlin_from_angles(angles, world_size, R, gf_params):  # initializations.py:80-116
    init_cells = perlin_from_angles(angles, world_size, R, gf_params)
    size = math.ceil(R) * 2
    pad_left = (128 - size) // 2
    pad_right = pad_left + 1 if pad_left * 2 + size != 128 else pad_left
    init_cells = np.pad(init_cells[:, :, 24:24 + size, 24:24 + size], ((0, 0), (0, 0), (pad_left, pad_right), (pad_left, pad_right)))
    return make_array_compressible(init_cells)


def random_uniform_from_unit(u: np.ndarray) -> np.ndarray:
    """initializations.py:10-32 after the random draw: ``u [nb_init, *world]`` uniform in [0, 1) -> states.
    ``jax.random.uniform(minval=0, maxval=m)`` is ``u * (m - 0) + 0`` clipped below at 0."""
    nb_init = u.shape[0]
    maxvals = np.linspace(0.4, 1., nb_init, dtype=np.float32).reshape((nb_init, ) + (1, ) * (u.ndim - 1))
    return make_array_compressible(np.asarray(u, np.float32) * maxvals)


# ---------------------------------------------------------------------------
# config + cell decoding (harness side)       leniax/utils.py, leniax/loader.py
# ---------------------------------------------------------------------------
def st2fracs2float(st: str) -> List[float]:  # utils.py:214-225
    return [float(Fraction(s)) for s in st.split(',')]


_OLD_GF = {0: 'poly_quad4', 1: 'gaussian', 2: 'gaussian_target', 3: 'step'}  # utils.py:102
_OLD_KF = {0: 'poly_quad', 1: 'gauss_bump', 2: 'step', 3: 'staircase', 4: 'gauss'}  # utils.py:103


def config_v1_to_v2(config: Dict) -> Dict:  # utils.py:91-150 (kernel part)
    new = []
    for kp in config['kernels_params']['k']:
        bs = st2fracs2float(kp['b']) if isinstance(kp['b'], str) else kp['b']
        new.append({
            'k_slug': 'circle_2d',
            'k_params': [kp['r'] if 'r' in kp else 1., bs],
            'kf_slug': _OLD_KF[kp['k_id']],
            'kf_params': [kp['q']],
            'gf_slug': _OLD_GF[kp['gf_id']],
            'gf_params': [kp['m'], kp['s']],
            'h': kp['h'],
            'c_in': kp['c_in'],
            'c_out': kp['c_out'],
        })
    config['kernels_params'] = new
    config['version'] = 2
    return config


def get_container(config: Dict) -> Dict:
    """utils.py:27-88 without Hydra/OmegaConf: fill world_size, scale, v1→v2."""
    config = copy.deepcopy(config)
    config.pop('hydra', None)
    rp, wp = config['render_params'], config['world_params']
    if rp.get('world_size', 'MISSING') == 'MISSING':
        rp['world_size'] = [2**rp['size_power2']] * wp['nb_dims']
    if rp.get('pixel_size', 'MISSING') == 'MISSING':
        rp['pixel_size'] = 2**rp.get('pixel_size_power2', 0)
    wp.setdefault('scale', 1.)
    if 'update_fn_version' in wp:
        wp['get_state_fn_slug'] = wp.pop('update_fn_version')
    if config.get('version', 1) == 1:
        config = config_v1_to_v2(config)
    return config


def load_yaml_config(path: str) -> Dict:
    import yaml
    with open(path, 'r', encoding='utf-8') as f:
        return get_container(yaml.safe_load(f))


def decompress_array_gzip(string_cells: str) -> np.ndarray:  # loader.py:105-129
    raw = gzip.decompress(base64.b64decode(string_cells))
    ints = np.frombuffer(raw, dtype='<i4')
    n = int(ints[0])
    vals, shape = ints[1:1 + n], [int(s) for s in ints[1 + n:]]
    return (vals.reshape(shape) / (NB_CHARS**2 - 1)).astype(np.float32)


def _rle_val(ch: str) -> int:  # loader.py:275-283
    if ch in '.b':
        return 0
    if ch == 'o':
        return 255
    if len(ch) == 1:
        return ord(ch) - ord('A') + 1
    return (ord(ch[0]) - ord('p')) * 24 + (ord(ch[1]) - ord('A') + 25)


def decompress_array_rle(code: str, nb_dims: int) -> np.ndarray:
    """Legacy run-length format of the test fixtures: loader.py:246-350."""
    delims = {'$': 1, '%': 2, '#': 3}
    closing = {1: '', 2: '$', 3: '%', 4: '#'}[nb_dims]
    stacks: List[List] = [[] for _ in range(nb_dims)]
    prefix, count = '', ''
    for ch in code.rstrip('!') + closing:
        if ch.isdigit():
            count += ch
        elif ch in 'pqrstuvwxy@':
            prefix = ch
        else:
            n = int(count) if count else 1
            if prefix + ch in delims:
                for d in range(delims[prefix + ch]):
                    stacks[d + 1].append(stacks[d])
                    stacks[d + 1].extend([[] for _ in range(n - 1)])
                    stacks[d] = []
            else:
                stacks[0].extend([_rle_val(prefix + ch) / 255] * n)
            prefix, count = '', ''
    nested = stacks[nb_dims - 1]

    lens = [0] * nb_dims

    def measure(d, lst):
        lens[d] = max(lens[d], len(lst))
        if d < nb_dims - 1:
            for sub in lst:
                measure(d + 1, sub)

    measure(0, nested)
    out = np.zeros(lens, dtype=np.float32)

    def fill(d, lst, idx):
        if d == nb_dims - 1:
            out[idx + (slice(0, len(lst)), )] = lst
        else:
            for i, sub in enumerate(lst):
                fill(d + 1, sub, idx + (i, ))

    fill(0, nested, ())
    return out


def load_raw_cells(config: Dict, use_init_cells: bool = True) -> np.ndarray:  # loader.py:206-240
    nb_dims = config['world_params']['nb_dims']
    rp = config['run_params']
    cells = rp['init_cells'] if (use_init_cells and 'init_cells' in rp) else rp['cells']
    if isinstance(cells, str):
        if cells == 'MISSING':
            cells = np.zeros((0, ), np.float32)
        else:
            try:
                cells = decompress_array_gzip(cells)
            except Exception:
                cells = decompress_array_rle(cells, nb_dims + 1)
    else:
        cells = np.array(cells, dtype=np.float32)
    if cells.ndim == nb_dims and config['world_params']['nb_channels'] == 1:
        cells = cells[np.newaxis]
    return cells.astype(np.float32)


def merge_cells(cells: np.ndarray, other: np.ndarray, offset=None) -> np.ndarray:  # utils.py:231-263
    assert cells.ndim == other.ndim and cells.shape[0] == other.shape[0]
    offset = offset or [0] * cells.ndim
    pads = []
    for i in range(cells.ndim):
        start = int(max((cells.shape[i] - other.shape[i]) // 2 + offset[i], 0))
        pads.append((start, int(cells.shape[i] - other.shape[i] - start)))
    return cells + np.pad(other, pads)


def init(config: Dict, use_init_cells: bool = True, fft: bool = True, dtype=np.float32):
    """helpers.py:35-88 → (cells [1,C,*dims], K, mapping).  Mutates config['world_params']['R'] when scale != 1, like the reference."""
    wp = config['world_params']
    world_size = list(config['render_params']['world_size'])
    raw_cells = load_raw_cells(config, use_init_cells)
    scale = wp.get('scale', 1.)
    if scale != 1.:  # helpers.py:58-66
        import scipy.ndimage
        raw_cells = np.array([scipy.ndimage.zoom(raw_cells[i], scale, order=0) for i in range(wp['nb_channels'])], dtype=np.float32)
        wp['R'] *= scale  # "the new R value will be used in the statistics" (helpers.py:65-66)
    if raw_cells.ndim > 1 + wp['nb_dims']:
        cells = raw_cells  # already [N, C, *dims]  (helpers.py:118-121)
    else:
        cells = np.zeros([wp['nb_channels']] + world_size, np.float32)
        if raw_cells.size:
            cells = merge_cells(cells, raw_cells)
        cells = cells[np.newaxis]
    K, mapping = get_kernels_and_mapping(config['kernels_params'], world_size, wp['nb_channels'], wp['R'], fft, dtype)
    return cells.astype(dtype), K, mapping


def init_and_run(config: Dict, use_init_cells: bool = True, with_jit: bool = True, fft: bool = True,
                 stat_trunc: bool = False, dtype=np.float32):
    """helpers.py:130-190."""
    config = copy.deepcopy(config)
    cells, K, mapping = init(config, use_init_cells, fft, dtype)
    wp = config['world_params']
    gf_params = mapping.get_gf_params(dtype)
    weights = mapping.get_kernels_weight_per_channel(dtype)
    update_fn = build_update_fn(mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), fft)
    stats_fn = build_compute_stats_fn(wp, config['render_params'], dtype)
    T = dtype(wp['T'])
    n_iter = config['run_params']['max_run_iter']
    if with_jit:
        return run_scan(cells, K, gf_params, weights, T, n_iter, update_fn, stats_fn)
    return run(cells, K, gf_params, weights, T, n_iter, update_fn, stats_fn, stat_trunc)
