"""Plan cache and launch glue between the Python entry points and the C ABI (include/leniax_b200.h).

PyTorch is used for device buffers and the current stream only; every arithmetic step of the path runs inside
``lnx_run_scan`` (csrc/lnx_kernels.cu).
"""
import ctypes
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import LNX_NB_STATS, STAT_KEYS, LnxDesc

STATE_FN_IDS = {'v1': 0, 'v2': 1, 'simple': 2}


def as_device_tensor(x, dtype, device) -> torch.Tensor:
    """Accept torch tensors (any device), numpy arrays or python numbers; return a contiguous tensor on ``device``."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    return t.to(device=device, dtype=dtype).contiguous()


def require_cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.LeniaxB200Error('leniax_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        raise _lib.LeniaxB200Error(f'leniax_b200 runs on CUDA devices only, got {device}')
    return device


class Plan:
    """Owns one ``lnx_plan`` (immutable description of the update + statistics functions)."""

    _cache: Dict[Tuple, 'Plan'] = {}

    def __init__(self, key: Tuple, desc: LnxDesc, device: torch.device):
        self.key = key
        self.desc = desc
        self.device = device
        self.lib = _lib.load_library()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.lnx_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
        self.handle = handle
        self.table_bytes = int(self.lib.lnx_kernel_table_bytes(handle))
        self.workspace_bytes = int(self.lib.lnx_workspace_bytes(handle))

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.lnx_plan_destroy(self.handle)
        except Exception:
            pass

    @classmethod
    def get(cls, *, world_size: Sequence[int], nb_channels: int, slots: Sequence[int], c_in: Sequence[int],
            gf_ids: Sequence[int], nb_slots: int, state_fn: str, weighted_average: bool, R: float, stats_dt: float,
            device: torch.device, force_tiled: bool = False, c_out: Optional[Sequence[int]] = None) -> 'Plan':
        c_out = tuple(c_out) if c_out is not None else tuple(_lib.LNX_COUT_ANY for _ in slots)
        key = (tuple(world_size), nb_channels, tuple(slots), tuple(c_in), tuple(gf_ids), nb_slots, state_fn,
               bool(weighted_average), float(R), float(stats_dt), str(device), c_out, bool(force_tiled))
        plan = cls._cache.get(key)
        if plan is None:
            if state_fn not in STATE_FN_IDS:
                raise NotImplementedError(f"state function '{state_fn}' is not one of {sorted(STATE_FN_IDS)}")
            if len(world_size) > 3:
                raise NotImplementedError('worlds with more than 3 dimensions are not supported')
            d = LnxDesc()
            d.nb_dims = len(world_size)
            for i, s in enumerate(world_size):
                d.dims[i] = int(s)
            d.nb_channels = nb_channels
            d.nb_kernels = len(slots)
            d.nb_slots = nb_slots
            if len(slots) > _lib.LNX_MAX_KERNELS:
                raise ValueError(f'at most {_lib.LNX_MAX_KERNELS} kernels are supported, got {len(slots)}')
            for k in range(len(slots)):
                d.slot[k], d.c_in[k], d.gf_id[k], d.c_out[k] = int(slots[k]), int(c_in[k]), int(gf_ids[k]), int(c_out[k])
            d.state_fn = STATE_FN_IDS[state_fn]
            d.weighted_average = 1 if weighted_average else 0
            d.R = float(R)
            d.stats_dt = float(stats_dt)
            d.flags = _lib.LNX_PLAN_FORCE_TILED if force_tiled else 0
            plan = cls(key, d, device)
            cls._cache[key] = plan
        return plan

    # ------------------------------------------------------------------------------------------------------------
    def prepare_kernels(self, K: torch.Tensor, n_sols: int) -> torch.Tensor:
        """K complex64 ``[n_sols, nb_slots, H, W]`` -> engine table (uint8 tensor of n_sols * table_bytes)."""
        K = K.to(device=self.device, dtype=torch.complex64).contiguous()
        table = torch.empty(n_sols * self.table_bytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.lnx_kernels_prepare(self.handle, n_sols, K.data_ptr(), table.data_ptr(), stream))
        return table

    def run_scan(self, cells0: torch.Tensor, K: torch.Tensor, gf_params: torch.Tensor, weights: torch.Tensor,
                 dt: torch.Tensor, max_run_iter: int, *, keep_trajectory: bool, flags: int = 0,
                 want_final_cells: bool = True) -> Dict[str, Optional[torch.Tensor]]:
        """cells0 ``[n_sols, n_init, C, *dims]``; K ``[n_sols, nb_slots, *dims]``; gf_params ``[n_sols, K, 2]``;
        weights ``[n_sols, C, K]``; dt ``[n_sols]``."""
        dev = self.device
        n_sols, n_init, C = cells0.shape[:3]
        dims = tuple(cells0.shape[3:])
        nk = self.desc.nb_kernels
        f32 = torch.float32
        table = self.prepare_kernels(K, n_sols)
        stats = torch.empty((LNX_NB_STATS, n_sols, max_run_iter, n_init), dtype=f32, device=dev)
        cm = torch.empty((n_sols, max_run_iter, n_init, C), dtype=f32, device=dev)
        if flags & _lib.LNX_RUN_EARLY_STOP:
            stats.zero_()
            cm.zero_()
        n_alive = torch.empty((n_sols, n_init), dtype=f32, device=dev)
        final = torch.empty_like(cells0) if want_final_cells else None
        traj_c = traj_f = traj_p = None
        if keep_trajectory:
            traj_c = torch.empty((n_sols, max_run_iter, n_init, C) + dims, dtype=f32, device=dev)
            traj_f = torch.empty_like(traj_c)
            traj_p = torch.empty((n_sols, max_run_iter, n_init, nk) + dims, dtype=f32, device=dev)
        ws = torch.empty(int(self.lib.lnx_workspace_bytes_for(self.handle, n_sols, n_init)), dtype=torch.uint8, device=dev)
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(
                self.lib.lnx_run_scan(
                    self.handle, n_sols, n_init, int(max_run_iter), int(flags), cells0.data_ptr(), table.data_ptr(),
                    gf_params.data_ptr(), weights.data_ptr(), dt.data_ptr(), stats.data_ptr(), cm.data_ptr(),
                    n_alive.data_ptr(), ptr(final), ptr(traj_c), ptr(traj_f), ptr(traj_p), ws.data_ptr(), ws.numel(), stream
                )
            )
        out = {k: stats[i] for i, k in enumerate(STAT_KEYS)}
        out['channel_mass'] = cm
        out['N'] = n_alive
        return {'stats': out, 'final_cells': final, 'cells': traj_c, 'field': traj_f, 'potential': traj_p}

    def update(self, state: torch.Tensor, K: torch.Tensor, gf_params: torch.Tensor, weights: torch.Tensor, dt: torch.Tensor
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One step (``lnx_update``): state ``[N, C, *dims]``, K ``[1, nb_slots, *dims]`` -> ``(state', field, potential)``."""
        dev = self.device
        N = state.shape[0]
        dims = tuple(state.shape[2:])
        table = self.prepare_kernels(K, 1)
        new_state, field = torch.empty_like(state), torch.empty_like(state)
        potential = torch.empty((N, self.desc.nb_kernels) + dims, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.lnx_update(self.handle, N, state.data_ptr(), table.data_ptr(), gf_params.data_ptr(), weights.data_ptr(),
                                           dt.data_ptr(), new_state.data_ptr(), field.data_ptr(), potential.data_ptr(), stream))
        return new_state, field, potential

    def variant(self, with_trajectory: bool) -> str:
        return self.lib.lnx_run_scan_variant(self.handle, 1 if with_trajectory else 0).decode()
