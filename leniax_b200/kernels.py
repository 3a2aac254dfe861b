"""Kernel rasterisation, packing and spectrum (reference: leniax/kernels.py:10-317).

``get_kernels_and_mapping`` keeps the reference signature and returns ``(K, KernelMapping)`` with the same shapes:
``K`` is ``complex64 [1, C, max_k_per_channel, H, W]`` for ``fft=True``.  On the GPU the kernels of ALL individuals of a
generation are rasterised by one launch (``lnx_rasterize_kernels``) and transformed by an exact small-support DFT
(``lnx_kernel_spectrum``; never cuFFT); the torch rasterisers below serve host-side callers (``fft=False`` on the CPU).
"""
import ctypes
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib
from .kernel_functions import register as kf_register


class KernelMapping(object):
    """Explicit mapping of the computation graph (leniax/kernels.py:10-63, same attributes)."""
    def __init__(self, nb_channels: int, nb_kernels: int):
        self.cin_kernels: List[List[int]] = [[] for _ in range(nb_channels)]
        self.cin_k_params: List[List] = [[] for _ in range(nb_channels)]
        self.cin_kfs: List[List[str]] = [[] for _ in range(nb_channels)]
        self.cin_gfs: List[List[str]] = [[] for _ in range(nb_channels)]
        self.cin_gf_params: List[List] = [[] for _ in range(nb_channels)]
        self.kernels_weight_per_channel: List[List[float]] = [[0.] * nb_kernels for _ in range(nb_channels)]
        self.true_channels: Optional[List[bool]] = None

    def get_k_params(self) -> torch.Tensor:
        return torch.tensor([p for sub in self.cin_k_params for p in sub], dtype=torch.float32)

    def get_gf_params(self, device=None) -> torch.Tensor:
        """``[nb_kernels, 2]`` float32 (kernels.py:48-55)."""
        return torch.tensor([p for sub in self.cin_gf_params for p in sub], dtype=torch.float32, device=device)

    def get_kernels_weight_per_channel(self, device=None) -> torch.Tensor:
        """``[nb_channels, nb_kernels]`` float32 (kernels.py:57-63)."""
        return torch.tensor(self.kernels_weight_per_channel, dtype=torch.float32, device=device)


def _default_device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')


def _grid(k_radius_px: int, scale: float, device) -> torch.Tensor:
    ax = torch.arange(2 * k_radius_px, device=device, dtype=torch.float32) - k_radius_px
    return torch.stack(torch.meshgrid(ax, ax, indexing='ij')) / scale  # [2, d0, d1]  (kernels.py:190-193)


def _shell(distances: torch.Tensor, bs: torch.Tensor, kf_slug: str, kf_params) -> torch.Tensor:
    nb_b = bs.shape[0]
    B_dist = nb_b * distances
    ring = bs[torch.clamp(torch.floor(B_dist).long(), max=nb_b - 1)]
    shell = kf_register[kf_slug](kf_params, torch.remainder(B_dist, 1))
    return (distances < 1).to(distances.dtype) * shell * ring  # kernels.py:199-207


def raw(R, k_params, kf_slug: str, kf_params, device=None) -> torch.Tensor:  # kernels.py:161-173
    if isinstance(k_params, torch.Tensor):
        return k_params.to(device=_default_device(device), dtype=torch.float32)
    return torch.tensor(k_params, dtype=torch.float32, device=_default_device(device))


def circle_2d(R, k_params: List, kf_slug: str, kf_params, device=None) -> torch.Tensor:  # kernels.py:176-212
    device = _default_device(device)
    r = k_params[0]
    bs = torch.tensor(k_params[1], dtype=torch.float32, device=device)
    cc = _grid(math.ceil(r * R), r * R, device)
    distances = torch.sqrt((cc**2).sum(dim=0))
    kernel = _shell(distances, bs, kf_slug, kf_params)
    return (kernel / kernel.sum())[None]


def _rotated(cc: torch.Tensor, theta: float) -> torch.Tensor:  # kernels.py:236-239
    c, s = math.cos(theta), math.sin(theta)
    return torch.stack([cc[0] * c + cc[1] * s, -cc[0] * s + cc[1] * c])


def ellipse_2d(R, k_params: List, kf_slug: str, kf_params, device=None) -> torch.Tensor:  # kernels.py:215-262
    device = _default_device(device)
    r, a, b, theta = k_params[0], k_params[2], k_params[3], k_params[4] * math.pi
    bs = torch.tensor(k_params[1], dtype=torch.float32, device=device)
    rc = _rotated(_grid(math.ceil(r * R), r * R, device), theta)
    distances = torch.sqrt((rc[0] / a)**2 + (rc[1] / b)**2)
    kernel = _shell(distances, bs, kf_slug, kf_params)
    kernel = kernel / kernel.sum()
    grad = rc[0].clone()
    grad[rc[0] < -0.01] = -1
    grad[grad > 0.01] = 1
    return (kernel * grad)[None]


def oriented_ellipse_2d(R, k_params: List, kf_slug: str, kf_params, device=None) -> torch.Tensor:  # kernels.py:265-309
    device = _default_device(device)
    r, a, b, theta = k_params[0], k_params[2], k_params[3], k_params[4] * math.pi
    bs = torch.tensor(k_params[1], dtype=torch.float32, device=device)
    rc = _rotated(_grid(math.ceil(r * R), r * R, device), theta)
    distances = torch.sqrt((rc[0] / a)**2 + (rc[1] / b)**2)
    kernel = _shell(distances, bs, kf_slug, kf_params) * rc[0]
    return (kernel / kernel.abs().sum())[None]


register: Dict[str, Callable] = {
    'raw': raw,
    'circle_2d': circle_2d,
    'ellipse_2d': ellipse_2d,
    'oriented_ellipse_2d': oriented_ellipse_2d,
}


def crop_zero(kernels: torch.Tensor) -> torch.Tensor:  # leniax/utils.py:296-318
    nz = kernels != 0
    if kernels.dim() == 3:
        return kernels[:, nz.any(dim=2).any(dim=0)][:, :, nz.any(dim=1).any(dim=0)]
    if kernels.dim() == 4:
        k1 = nz.flatten(2).any(dim=2).any(dim=0)
        k2 = nz.any(dim=3).any(dim=1).any(dim=0)
        k3 = nz.any(dim=2).any(dim=1).any(dim=0)
        return kernels[:, k1][:, :, k2][:, :, :, k3]
    raise ValueError("Can't handle more than 3 dimensions")


def rfftn_full(images: torch.Tensor, nb_dims: int) -> torch.Tensor:
    """``fftn`` over the last ``nb_dims`` axes of real float32 CUDA images with the engine's own FFT (``lnx_rfftn``:
    resident 128x128 butterflies or the tiled multi-pass engine; never cuFFT)."""
    if not images.is_cuda:
        raise _lib.LeniaxB200Error('the kernel spectrum is computed on the GPU by lnx_rfftn; no CPU fallback exists')
    lib = _lib.load_library()
    dims = tuple(images.shape[-nb_dims:])
    flat = images.reshape((-1, ) + dims).contiguous().float()
    out = torch.empty(flat.shape, dtype=torch.complex64, device=flat.device)
    cdims = (ctypes.c_int32 * 3)(*(list(dims) + [1] * (3 - nb_dims)))
    with torch.cuda.device(flat.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.lnx_rfftn(nb_dims, cdims, flat.shape[0], flat.data_ptr(), out.data_ptr(), stream))
    return out.reshape(images.shape)


def rfft2_full(images: torch.Tensor) -> torch.Tensor:
    return rfftn_full(images, 2)


def sphere_nd(R, k_params: List, kf_slug: str, kf_params, device=None, nb_dims: int = 3) -> torch.Tensor:
    """EXTENSION (not in the reference, which only ships ``*_2d`` generators, kernels.py:312-317): the ``circle_2d``
    formula evaluated with an ``nb_dims``-dimensional distance.  Used to build the ``raw`` 3-D kernels of BASELINE
    config E; the result is a plain array, i.e. what ``k_slug: raw`` carries."""
    device = _default_device(device)
    r = k_params[0]
    bs = torch.tensor(k_params[1], dtype=torch.float32, device=device)
    k_radius_px = math.ceil(r * R)
    ax = (torch.arange(2 * k_radius_px, device=device, dtype=torch.float32) - k_radius_px) / (r * R)
    grids = torch.meshgrid(*([ax] * nb_dims), indexing='ij')
    distances = torch.sqrt(sum(g**2 for g in grids))
    kernel = _shell(distances, bs, kf_slug, kf_params)
    return (kernel / kernel.sum())[None]


def _fill_mapping(kernels_params: List, nb_channels: int) -> KernelMapping:
    """kernels.py:90-143: sort IN PLACE by ``c_in``, record the computation graph, mark the padded kernel slots."""
    mapping = KernelMapping(nb_channels, len(kernels_params))
    kernels_params.sort(key=lambda d: d['c_in'])
    for idx, p in enumerate(kernels_params):
        mapping.cin_kernels[p['c_in']].append(idx)
        mapping.cin_gfs[p['c_in']].append(p['gf_slug'])
        mapping.cin_gf_params[p['c_in']].append(p['gf_params'])
        mapping.cin_kfs[p['c_in']].append(p['kf_slug'])
        mapping.cin_k_params[p['c_in']].append(p['k_params'])
        mapping.kernels_weight_per_channel[p['c_out']][idx] = p['h']
    max_k = max(len(lst) for lst in mapping.cin_kernels)
    true_channels: List[bool] = []
    for lst in mapping.cin_kernels:  # kernels.py:122-143
        true_channels += [True] * len(lst) + [False] * (max_k - len(lst))
    mapping.true_channels = None if all(true_channels) else true_channels
    return mapping


def _spec_of(p: Dict, R: float) -> Tuple[Optional[_lib.LnxKernelSpec], int]:
    """Parametric kernel -> (descriptor for ``lnx_rasterize_kernels``, radius in pixels); ``raw`` kernels -> ``(None, 0)``."""
    slug = p['k_slug']
    if slug == 'raw':
        return None, 0
    if slug not in _lib.KSHAPE_IDS:
        raise NotImplementedError(f"kernel shape '{slug}' is not one of {sorted(register)}")
    if p['kf_slug'] not in _lib.KF_IDS:
        raise NotImplementedError(f"kernel function '{p['kf_slug']}' is not one of {sorted(_lib.KF_IDS)}")
    kp = p['k_params']
    bs = [float(b) for b in (kp[1].tolist() if hasattr(kp[1], 'tolist') else kp[1])]
    if len(bs) > _lib.LNX_MAX_RINGS:
        raise ValueError(f'at most {_lib.LNX_MAX_RINGS} rings per kernel are supported, got {len(bs)}')
    spec = _lib.LnxKernelSpec()
    spec.shape, spec.kf, spec.nb_b, spec.r = _lib.KSHAPE_IDS[slug], _lib.KF_IDS[p['kf_slug']], len(bs), float(kp[0])
    for i, b in enumerate(bs):
        spec.bs[i] = b
    kfp = [float(v) for v in (p['kf_params'].tolist() if hasattr(p['kf_params'], 'tolist') else p['kf_params'])]
    for i, v in enumerate(kfp[:2]):
        spec.kf_params[i] = v
    if slug != 'circle_2d':
        theta = float(kp[4]) * math.pi
        spec.a, spec.b, spec.cos_theta, spec.sin_theta = float(kp[2]), float(kp[3]), math.cos(theta), math.sin(theta)
    return spec, math.ceil(float(kp[0]) * R)


def _spatial_kernels_cuda(all_params: List[Dict], R: float, device) -> List[torch.Tensor]:
    """Spatial kernels ``[*support]`` of a flat list of kernel descriptions, all parametric ones rasterised by ONE launch."""
    specs = [_spec_of(p, R) for p in all_params]
    side = 2 * max([k for _, k in specs] + [1])
    n_par = sum(1 for sp, _ in specs if sp is not None)
    out: List[Optional[torch.Tensor]] = [None] * len(all_params)
    if n_par:
        arr = (_lib.LnxKernelSpec * n_par)(*[sp for sp, _ in specs if sp is not None])
        buf = torch.empty((n_par, side, side), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(_lib.load_library().lnx_rasterize_kernels(n_par, arr, float(R), side, buf.data_ptr(), torch.cuda.current_stream().cuda_stream))
        j = 0
        for i, (sp, _) in enumerate(specs):
            if sp is not None:
                out[i] = buf[j]
                j += 1
    for i, p in enumerate(all_params):
        if out[i] is None:
            out[i] = raw(R, p['k_params'], p['kf_slug'], p['kf_params'], device=device)[0]
    return out  # type: ignore


def kernel_spectrum(spatial: torch.Tensor, world_size: List[int]) -> torch.Tensor:
    """``fftn(fftshift(centre-pad(kernel)))`` (kernels.py:145-149, utils.py:231-263) of ``spatial [n, *support]`` -> complex64
    ``[n, *world_size]``: exact separable DFT over the support (``lnx_kernel_spectrum``; fp64 inside, no cuFFT)."""
    if not spatial.is_cuda:
        raise _lib.LeniaxB200Error('the kernel spectrum is computed on the GPU by lnx_kernel_spectrum; no CPU fallback exists')
    nd = len(world_size)
    spatial = spatial.contiguous().float()
    n = spatial.shape[0]
    K = torch.empty((n, ) + tuple(world_size), dtype=torch.complex64, device=spatial.device)
    dims = (ctypes.c_int32 * 3)(*(list(world_size) + [1] * (3 - nd)))
    support = (ctypes.c_int32 * 3)(*(list(spatial.shape[1:]) + [1] * (3 - nd)))
    with torch.cuda.device(spatial.device):
        _lib.check(_lib.load_library().lnx_kernel_spectrum(nd, dims, n, support, spatial.data_ptr(), K.data_ptr(),
                                                           torch.cuda.current_stream().cuda_stream))
    return K


def _spectra_of(spatials: List[torch.Tensor], world_size: List[int]) -> torch.Tensor:
    """Spectra of a list of spatial kernels (grouped by support so that every group is one batched call) -> ``[n, *world_size]``."""
    out = torch.empty((len(spatials), ) + tuple(world_size), dtype=torch.complex64, device=spatials[0].device)
    groups: Dict[Tuple[int, ...], List[int]] = {}
    for i, k in enumerate(spatials):
        groups.setdefault(tuple(k.shape), []).append(i)
    for shape, idxs in groups.items():
        if any(s > w for s, w in zip(shape, world_size)):
            raise ValueError(f'kernel of size {shape} does not fit the world {tuple(world_size)}')
        out[torch.tensor(idxs, device=out.device)] = kernel_spectrum(torch.stack([spatials[i] for i in idxs]), world_size)
    return out


def get_kernels_and_mapping_batch(all_kernels_params: List[List], world_size: List[int], nb_channels: int, R: float, device=None
                                  ) -> Tuple[torch.Tensor, List[KernelMapping]]:
    """``get_kernels_and_mapping(fft=True)`` for every individual of a QD generation at once (replaces the per-individual loop of
    leniax/qd.py:113-131): ONE rasterisation launch and ONE spectrum call (2-3 launches) for all kernels of all individuals.
    Returns ``K [n_sols, 1, C, max_k, *world_size]`` complex64 and the mappings.  Every ``kernels_params`` is sorted in place."""
    device = _default_device(device)
    world_size = list(world_size)
    mappings = [_fill_mapping(kp, nb_channels) for kp in all_kernels_params]
    max_k = max(len(lst) for lst in mappings[0].cin_kernels)
    layout = [list(m.cin_kernels) for m in mappings]
    if any(lay != layout[0] for lay in layout):
        raise ValueError('get_kernels_and_mapping_batch: all individuals must share the same (c_in -> kernels) layout')
    flat = [p for kp in all_kernels_params for p in kp]
    spectra = _spectra_of(_spatial_kernels_cuda(flat, R, device), world_size)
    n_sols, nk = len(all_kernels_params), len(all_kernels_params[0])
    spectra = spectra.reshape((n_sols, nk) + tuple(world_size))
    K = torch.zeros((n_sols, nb_channels, max_k) + tuple(world_size), dtype=torch.complex64, device=device)
    for c, lst in enumerate(mappings[0].cin_kernels):
        for j, idx in enumerate(lst):
            K[:, c, j] = spectra[:, idx]
    return K[:, None], mappings


def get_kernels_and_mapping(kernels_params: List, world_size: List[int], nb_channels: int, R: float, fft: bool = True,
                            device=None) -> Tuple[torch.Tensor, KernelMapping]:
    """Construct the kernel array and the associated mapping (leniax/kernels.py:66-158).

    Like the reference it sorts ``kernels_params`` **in place** by ``c_in`` (kernels.py:90).  On a CUDA device the kernels are
    rasterised by ``lnx_rasterize_kernels`` and, for ``fft=True``, transformed by ``lnx_kernel_spectrum``; the torch
    rasterisers of this module serve host-side callers of the direct-convolution path (``fft=False`` on the CPU).
    """
    device = _default_device(device)
    world_size = list(world_size)
    if fft and device.type == 'cuda':
        # K only depends on the kernel shapes (not on growth parameters or weights): callers that loop over individuals which
        # share them (conf/config_qd_cmame*.yaml mutate gf_params and h) get the cached spectrum back.
        mapping = _fill_mapping(kernels_params, nb_channels)
        cache_key = _kernel_cache_key(kernels_params, world_size, nb_channels, R, fft, device)
        cached = _K_CACHE.get(cache_key) if cache_key is not None else None
        if cached is not None:
            return cached.clone(), mapping
        K = get_kernels_and_mapping_batch([kernels_params], world_size, nb_channels, R, device)[0][0]
        if cache_key is not None and _K_CACHE_MAX > 0:
            if len(_K_CACHE) >= _K_CACHE_MAX:
                _K_CACHE.pop(next(iter(_K_CACHE)))
            _K_CACHE[cache_key] = K.clone()
        return K, mapping
    if fft:
        raise _lib.LeniaxB200Error('the kernel spectrum is computed on the GPU (lnx_kernel_spectrum); no CPU fallback exists')
    mapping = _fill_mapping(kernels_params, nb_channels)
    max_k = max(len(lst) for lst in mapping.cin_kernels)
    if device.type == 'cuda':
        ks = [k[None] for k in _spatial_kernels_cuda(kernels_params, R, device)]
    else:
        ks = [register[p['k_slug']](R, p['k_params'], p['kf_slug'], p['kf_params'], device=device) for p in kernels_params]
    padded = []
    for k in ks:
        pads: List[int] = []
        for ws, ksz in reversed(list(zip(world_size, k.shape[1:]))):  # F.pad wants the last dim first
            lo = (ws - ksz) // 2
            pads += [lo, lo if (ws - ksz) % 2 == 0 else lo + 1]
        padded.append(torch.nn.functional.pad(k, pads))
    kernels = crop_zero(torch.cat(padded))  # [nb_kernels, kh, kw]
    kshape = tuple(kernels.shape[1:])
    per_channel = []
    for lst in mapping.cin_kernels:
        kc = kernels[torch.tensor(lst, dtype=torch.long, device=device)] if lst else kernels.new_zeros((0, ) + kshape)
        missing = max_k - kc.shape[0]
        if missing:
            kc = torch.cat([kc, kernels.new_zeros((missing, ) + kshape)])
        per_channel.append(kc)
    return torch.cat(per_channel)[:, None], mapping  # [C*max_k, 1, kh, kw]


def is_pow2_world(world_size) -> bool:
    return all(int(n) >= 8 and (int(n) & (int(n) - 1)) == 0 for n in world_size)


def spatial_from_spectrum(K_fft: torch.Tensor, nb_slots: int, world_size) -> torch.Tensor:
    """The kernels of the direct-convolution path, ``[nb_slots, 1, kh, kw]`` float32 (the layout of ``get_kernels_and_mapping(fft=False)``,
    kernels.py:116-117), recovered from the FFT kernels ``K = fftn(fftshift(centre-pad(kernel)))`` of a 2-D world of ANY size.

    Worlds whose size is not a power of two (the reference's ``fftn`` takes any size, core.py:81) are stepped through
    ``lnx_update_conv``: ``real(ifftn(K))`` is the circular-convolution kernel ``k[u]`` the FFT path applies, ``potential[y] = sum_u
    state[y - u] k[u]``; the cross-correlation taps ``lnx_update_conv`` wants are ``taps[i] = k[(h - i) mod N]`` with the centre at ``h =
    kh // 2``.  The inverse transform is two dense products with the DFT matrix in complex128 (a set-up step, once per scan); entries
    below 2e-6 of the largest tap (ten times the rounding noise a complex64 spectrum leaves on the taps) are outside the measured support."""
    if len(world_size) != 2:
        raise NotImplementedError('worlds whose size is not a power of two are supported in 2-D only (direct-convolution potential, as in the '
                                  'reference: core.py:136)')
    H, W = int(world_size[0]), int(world_size[1])
    K = K_fft.reshape(nb_slots, H, W).to(torch.complex128)
    dev = K.device

    def idft(n):
        a = torch.arange(n, device=dev, dtype=torch.float64)
        ang = 2 * torch.pi * torch.outer(a, a) / n
        return torch.complex(torch.cos(ang), torch.sin(ang)) / n

    k = (idft(H) @ K @ idft(W)).real  # [nb_slots, H, W]: k[u, v], the origin at index (0, 0)
    mag = k.abs().amax(dim=0)
    live = mag > 2e-6 * float(mag.max())
    if not bool(live.any()):
        return torch.zeros((nb_slots, 1, 1, 1), dtype=torch.float32, device=dev)
    uy = torch.arange(H, device=dev)
    ux = torch.arange(W, device=dev)
    dy = torch.minimum(uy, H - uy)  # circular distance from the origin
    dx = torch.minimum(ux, W - ux)
    hy = int(dy[live.any(dim=1)].max())
    hx = int(dx[live.any(dim=0)].max())
    hy, hx = min(hy, (H - 1) // 2), min(hx, (W - 1) // 2)
    iy = (hy - torch.arange(2 * hy + 1, device=dev)) % H  # taps[i] = k[(h - i) mod N]
    ix = (hx - torch.arange(2 * hx + 1, device=dev)) % W
    taps = k[:, iy][:, :, ix]
    return taps.to(torch.float32)[:, None].contiguous()


_K_CACHE: Dict[Tuple, torch.Tensor] = {}
_K_CACHE_MAX = 64
_UNCACHEABLE = object()


def _freeze(x):
    """Hashable image of a kernel parameter; ``_UNCACHEABLE`` for arrays (``k_slug: raw``) at any nesting level."""
    import numbers
    if isinstance(x, (list, tuple)):
        items = tuple(_freeze(v) for v in x)
        return _UNCACHEABLE if any(v is _UNCACHEABLE for v in items) else items
    if isinstance(x, (str, bool)) or x is None:
        return x
    if isinstance(x, numbers.Real):  # Python and NumPy scalars
        return float(x)
    if hasattr(x, 'ndim') and x.ndim == 0 and hasattr(x, 'item'):  # 0-d arrays / tensors
        return float(x.item())
    return _UNCACHEABLE


def _kernel_cache_key(kernels_params: List, world_size: List[int], nb_channels: int, R: float, fft: bool, device) -> Optional[Tuple]:
    parts = []
    for p in kernels_params:
        kp, kfp = _freeze(p['k_params']), _freeze(p['kf_params'])
        if kp is _UNCACHEABLE or kfp is _UNCACHEABLE:
            return None
        parts.append((p['k_slug'], kp, p['kf_slug'], kfp, int(p['c_in'])))
    return (tuple(parts), tuple(world_size), int(nb_channels), float(R), bool(fft), str(torch.device(device)))
