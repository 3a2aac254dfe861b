"""Initial-state generators for the batched search (reference: leniax/initializations.py:10-123, leniax/perlin.py:16-71).

Generated on the device that will run the simulation (no 268 MB host->device copy per generation) by hand-written launches
(``lnx_random_uniform`` / ``lnx_init_perlin`` / ``lnx_init_uniform``, csrc/lnx_tu_setup.cu): one CTA per world, all worlds of a
call — or, through ``perlin_batch``, of a whole QD generation — in one launch.  Random numbers come from a counter-based key
(``RngKey``, SplitMix64 of seed + index); bit-parity with ``jax.random`` (threefry) is NOT provided — no reference test pins a
random draw (SURVEY.md §8c "parity unpinned").  With ``device='cpu'`` the same functions run as torch ops (host-side tooling and
the CPU tests; the scan itself has no CPU path).
"""
import ctypes
import math
from typing import Callable, Dict, List, Sequence, Tuple

import torch

from . import _lib
from .loader import make_array_compressible


class RngKey:
    """Minimal splittable key (stands in for ``jax.random.PRNGKey``): a 64-bit integer, split with SplitMix64."""
    _MASK = (1 << 64) - 1

    def __init__(self, seed: int):
        self.seed = int(seed) & self._MASK

    @staticmethod
    def _mix(z: int) -> int:
        z = (z + 0x9E3779B97F4A7C15) & RngKey._MASK
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & RngKey._MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & RngKey._MASK
        return z ^ (z >> 31)

    def split(self, num: int = 2) -> List['RngKey']:
        return [RngKey(self._mix(self.seed + 0x632BE59BD9B4E019 * (i + 1))) for i in range(num)]

    def generator(self, device) -> torch.Generator:
        g = torch.Generator(device=device)
        g.manual_seed(self.seed & ((1 << 63) - 1))
        return g

    def tolist(self):
        return [self.seed >> 32, self.seed & 0xFFFFFFFF]

    def __repr__(self):
        return f'RngKey({self.seed:#x})'


def _device(device=None):
    if device is not None:
        return torch.device(device)
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')


def interpolant(t):
    return t * t * t * (t * (t * 6 - 15) + 10)


def generate_perlin_noise_2d(angles: torch.Tensor, shape: Tuple[int, int], res: Tuple[int, int], nb_noise: int = 1) -> torch.Tensor:
    """leniax/perlin.py:16-71, including its ``diff``-based corner slicing."""
    if angles.is_cuda:
        return _perlin_cuda(angles, list(shape), noise_only=True)
    gradients = torch.stack([torch.cos(angles), torch.sin(angles)], dim=-1)  # [n, Hr, Wr, 2]
    gradients = torch.cat([gradients, gradients[:, :1]], dim=1)
    gradients = torch.cat([gradients, gradients[:, :, :1]], dim=2)  # wrap pad -> [n, Hr+1, Wr+1, 2]
    d = (shape[0] // res[0], shape[1] // res[1])
    gradients = gradients.repeat_interleave(d[0], dim=1).repeat_interleave(d[1], dim=2)
    diff = [gradients.shape[1] - shape[0], gradients.shape[2] - shape[1]]
    g00 = gradients[:, :-diff[0], :-diff[1]]
    g10 = gradients[:, diff[0]:, :-diff[1]]
    g01 = gradients[:, :-diff[0], diff[1]:]
    g11 = gradients[:, diff[0]:, diff[1]:]
    dev = angles.device
    gy = (torch.arange(shape[0], device=dev, dtype=torch.float32) * (res[0] / shape[0])) % 1
    gx = (torch.arange(shape[1], device=dev, dtype=torch.float32) * (res[1] / shape[1])) % 1
    g0, g1 = torch.meshgrid(gy, gx, indexing='ij')
    g0, g1 = g0[None], g1[None]
    n00 = g0 * g00[..., 0] + g1 * g00[..., 1]
    n10 = (g0 - 1) * g10[..., 0] + g1 * g10[..., 1]
    n01 = g0 * g01[..., 0] + (g1 - 1) * g01[..., 1]
    n11 = (g0 - 1) * g11[..., 0] + (g1 - 1) * g11[..., 1]
    t0, t1 = interpolant(g0), interpolant(g1)
    n0 = n00 * (1 - t0) + t0 * n10
    n1 = n01 * (1 - t0) + t0 * n11
    return math.sqrt(2) * ((1 - t1) * n0 + t1 * n1)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _uniform01(key: RngKey, shape: Sequence[int], device: torch.device) -> torch.Tensor:
    """Uniform [0, 1) numbers of ``key``: ``lnx_random_uniform`` on a GPU, ``torch.Generator`` on the host."""
    if device.type != 'cuda':
        return torch.rand(list(shape), generator=key.generator(device), device=device)
    out = torch.empty(list(shape), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().lnx_random_uniform(ctypes.c_uint64(key.seed), out.numel(), out.data_ptr(), _stream(device)))
    return out


def random_uniform(rng_key: RngKey, nb_init: int, world_size: List[int], R: float, gf_params: List, device=None):
    """initializations.py:10-32 (maxvals broadcast over the world axes)."""
    device = _device(device)
    rng_key, subkey = rng_key.split()
    maxvals = torch.linspace(0.4, 1., nb_init, device=device)
    if device.type != 'cuda':
        cells = torch.rand([nb_init] + list(world_size), generator=subkey.generator(device), device=device)
        return rng_key, make_array_compressible(cells * maxvals.reshape([nb_init] + [1] * len(world_size)))
    cells = torch.empty([nb_init] + list(world_size), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().lnx_init_uniform(ctypes.c_uint64(subkey.seed), nb_init, math.prod(world_size), maxvals.data_ptr(),
                                                        cells.data_ptr(), _stream(device)))
    return rng_key, cells


def perlin_resolution(world_size: List[int], R: float) -> List[int]:
    kernel_radius = math.ceil(R)
    return [world_size[0] // (kernel_radius * 3), world_size[1] // (kernel_radius * 2)]  # initializations.py:56


def perlin_scaling(nb_init: int, gf_params: List) -> List[float]:
    lo = gf_params[0]
    hi = min(1, 3 * lo)  # initializations.py:57-58
    return [lo + i / nb_init * (hi - lo) for i in range(nb_init)]


def _perlin_cuda(angles: torch.Tensor, world_size: List[int], scaling: torch.Tensor = None, noise_only: bool = False) -> torch.Tensor:
    angles = angles.contiguous().float()
    n, res0, res1 = angles.shape
    out = torch.empty((n, world_size[0], world_size[1]), dtype=torch.float32, device=angles.device)
    with torch.cuda.device(angles.device):
        _lib.check(_lib.load_library().lnx_init_perlin(n, world_size[0], world_size[1], res0, res1, angles.data_ptr(),
                                                       None if noise_only else scaling.data_ptr(), None if noise_only else out.data_ptr(),
                                                       out.data_ptr() if noise_only else None, _stream(angles.device)))
    return out


def perlin_from_angles(angles: torch.Tensor, world_size: List[int], R: float, gf_params: List) -> torch.Tensor:
    """initializations.py:56-75 after the random draw: ``angles [nb_init, res0, res1]`` -> states ``[nb_init, 1, H, W]``."""
    nb_init = angles.shape[0]
    scaling = torch.tensor(perlin_scaling(nb_init, gf_params), dtype=torch.float32, device=angles.device)
    if angles.is_cuda:
        return _perlin_cuda(angles, world_size, scaling)[:, None]
    res = perlin_resolution(world_size, R)
    cells = generate_perlin_noise_2d(angles, tuple(world_size), tuple(res), nb_init)
    cells = cells - cells.amin(dim=(1, 2), keepdim=True)
    cells = cells / cells.amax(dim=(1, 2), keepdim=True)
    cells = cells * scaling[:, None, None]
    return make_array_compressible(cells[:, None])


def perlin(rng_key: RngKey, nb_init: int, world_size: List[int], R: float, gf_params: List, device=None):
    """initializations.py:35-77."""
    device = _device(device)
    res = perlin_resolution(world_size, R)
    rng_key, subkey = rng_key.split()
    angles = 2 * math.pi * _uniform01(subkey, [nb_init] + res, device)
    return rng_key, perlin_from_angles(angles, world_size, R, gf_params)


def perlin_batch(rng_keys: List[RngKey], nb_init: int, world_size: List[int], R: float, all_gf_params: List[List], device=None):
    """``perlin`` for every individual of a QD generation (its own key and growth parameters each; replaces the per-individual calls of
    leniax/qd.py:121-125): same states as ``len(rng_keys)`` separate ``perlin`` calls, generated by ONE launch.
    Returns ``(new keys, cells [n_sols, nb_init, 1, H, W])``."""
    device = _device(device)
    if device.type != 'cuda':
        outs = [perlin(k, nb_init, world_size, R, g, device) for k, g in zip(rng_keys, all_gf_params)]
        return [o[0] for o in outs], torch.stack([o[1] for o in outs])
    res = perlin_resolution(world_size, R)
    splits = [k.split() for k in rng_keys]
    seeds = (ctypes.c_uint64 * len(splits))(*[sub.seed for _, sub in splits])
    scaling = torch.tensor([v for g in all_gf_params for v in perlin_scaling(nb_init, g)], dtype=torch.float32, device=device)
    cells = torch.empty((len(splits), nb_init, 1, world_size[0], world_size[1]), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().lnx_init_perlin_seeded(len(splits), seeds, nb_init, world_size[0], world_size[1], res[0], res[1],
                                                              scaling.data_ptr(), cells.data_ptr(), _stream(device)))
    return [nk for nk, _ in splits], cells


def cropped_perlin(rng_key: RngKey, nb_init: int, world_size: List[int], R: float, gf_params: List, device=None):
    """initializations.py:80-116."""
    rng_key, init_cells = perlin(rng_key, nb_init, world_size, R, gf_params, device)
    size = math.ceil(R) * 2
    pad_left = (128 - size) // 2
    pad_right = pad_left + 1 if pad_left * 2 + size != 128 else pad_left
    init_cells = torch.nn.functional.pad(init_cells[:, :, 24:24 + size, 24:24 + size], (pad_left, pad_right, pad_left, pad_right))
    return rng_key, make_array_compressible(init_cells)


register: Dict[str, Callable] = {
    'random_uniform': random_uniform,
    'perlin': perlin,
    'cropped_perlin': cropped_perlin,
}
