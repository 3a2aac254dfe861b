"""Kernel rasterisation, packing and spectrum (reference: leniax/kernels.py:10-317).

``get_kernels_and_mapping`` keeps the reference signature and returns ``(K, KernelMapping)`` with the same shapes:
``K`` is ``complex64 [1, C, max_k_per_channel, H, W]`` for ``fft=True``.  The spectrum is computed on the GPU with the
engine's own butterflies (``lnx_rfft2``), not cuFFT; rasterisation uses torch elementwise ops on the same device.
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


def get_kernels_and_mapping(kernels_params: List, world_size: List[int], nb_channels: int, R: float, fft: bool = True,
                            device=None) -> Tuple[torch.Tensor, KernelMapping]:
    """Construct the kernel array and the associated mapping (leniax/kernels.py:66-158).

    Like the reference it sorts ``kernels_params`` **in place** by ``c_in`` (kernels.py:90).
    """
    device = _default_device(device)
    world_size = list(world_size)
    mapping = KernelMapping(nb_channels, len(kernels_params))
    kernels_params.sort(key=lambda d: d['c_in'])
    # K only depends on the kernel shapes (not on growth parameters or weights): in a QD generation every individual usually
    # shares them (conf/config_qd_cmame*.yaml mutate gf_params and h), so the rasterisation + FFT is done once and reused.
    cache_key = _kernel_cache_key(kernels_params, world_size, nb_channels, R, fft, device)
    cached = _K_CACHE.get(cache_key) if cache_key is not None else None
    padded = []
    for idx, p in enumerate(kernels_params):
        if cached is None:
            k = register[p['k_slug']](R, p['k_params'], p['kf_slug'], p['kf_params'], device=device)
            pads: List[int] = []
            for ws, ks in reversed(list(zip(world_size, k.shape[1:]))):  # F.pad wants the last dim first
                lo = (ws - ks) // 2
                pads += [lo, lo if (ws - ks) % 2 == 0 else lo + 1]
            padded.append(torch.nn.functional.pad(k, pads))
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
    if cached is not None:
        return cached.clone(), mapping

    kernels = torch.cat(padded)  # [nb_kernels, *dims]
    if not fft:
        kernels = crop_zero(kernels)
    kshape = tuple(kernels.shape[1:])
    per_channel = []
    for lst in mapping.cin_kernels:
        kc = kernels[torch.tensor(lst, dtype=torch.long, device=device)] if lst else kernels.new_zeros((0, ) + kshape)
        missing = max_k - kc.shape[0]
        if missing:
            kc = torch.cat([kc, kernels.new_zeros((missing, ) + kshape)])
        per_channel.append(kc)

    if fft:
        nd = len(world_size)
        dims = tuple(range(-nd, 0))
        K = torch.stack(per_channel)[None]  # [1, C, max_k, *dims]
        K = torch.roll(K, shifts=[s // 2 for s in K.shape[-nd:]], dims=dims)  # fftshift (kernels.py:147)
        K = rfftn_full(K, nd)  # kernels.py:148
    else:
        K = torch.cat(per_channel)[:, None]  # [C*max_k, 1, kh, kw]
    if cache_key is not None and _K_CACHE_MAX > 0:
        if len(_K_CACHE) >= _K_CACHE_MAX:
            _K_CACHE.pop(next(iter(_K_CACHE)))
        _K_CACHE[cache_key] = K.clone()
    return K, mapping


_K_CACHE: Dict[Tuple, torch.Tensor] = {}
_K_CACHE_MAX = 64


def _freeze(x):
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(v) for v in x)
    if isinstance(x, (int, float, str, bool)) or x is None:
        return x
    return None  # tensors / arrays (k_slug 'raw'): not cached


def _kernel_cache_key(kernels_params: List, world_size: List[int], nb_channels: int, R: float, fft: bool, device) -> Optional[Tuple]:
    parts = []
    for p in kernels_params:
        kp, kfp = _freeze(p['k_params']), _freeze(p['kf_params'])
        if kp is None or kfp is None or (isinstance(kp, tuple) and any(v is None for v in kp)):
            return None
        parts.append((p['k_slug'], kp, p['kf_slug'], kfp, int(p['c_in'])))
    return (tuple(parts), tuple(world_size), int(nb_channels), float(R), bool(fft), str(torch.device(device)))
