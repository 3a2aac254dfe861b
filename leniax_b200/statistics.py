"""Statistics descriptors and stop criteria (reference: leniax/statistics.py).

``build_compute_stats_fn`` returns a descriptor (R, dt of the build-time config, world size) instead of a traced
closure; the 12 statistics, their carry and ``check_heuristics`` are computed inside the persistent kernel
(csrc/lnx_step.cuh: ``cells_fused`` / ``stats_finalize``).
"""
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from .constant import EPSILON

MONOTONIC_STOP_STEP = 128  # statistics.py:284
MASS_VOLUME_THRESHOLD = 10.  # statistics.py:313
MASS_VOLUME_STOP_STEP = 128  # statistics.py:314


@dataclass(frozen=True)
class ComputeStatsFn:
    """Closure constants of ``compute_stats`` (statistics.py:22-33)."""
    world_size: Tuple[int, ...]
    R: float
    dt: float

    def __call__(self, cells, field, potential, total_shift_idx, mass_centroid, mass_angle):
        """Stand-alone ``compute_stats`` (statistics.py:36-126): ``cells/field [N, C, *dims]``, ``potential [N, K, *dims]``,
        ``total_shift_idx [N, D]`` int32, ``mass_centroid [D, N]``, ``mass_angle [N]`` ->
        ``(stats dict, total_shift_idx, mass_centroid, mass_angle)``.  Inside the scans the same statistics are fused
        into the persistent kernels; this entry point runs ``lnx_compute_stats``."""
        from . import _lib, engine
        dev = engine.require_cuda_device(cells.device if isinstance(cells, torch.Tensor) and cells.is_cuda else None)
        f32 = torch.float32
        cells = engine.as_device_tensor(cells, f32, dev)
        field = engine.as_device_tensor(field, f32, dev)
        potential = engine.as_device_tensor(potential, f32, dev)
        N, C, K = cells.shape[0], cells.shape[1], potential.shape[1]
        nd = len(self.world_size)
        if tuple(cells.shape[2:]) != tuple(self.world_size):
            raise ValueError(f'compute_stats_fn was built for world_size {self.world_size}, cells are {tuple(cells.shape[2:])}')
        shift = engine.as_device_tensor(total_shift_idx, torch.int32, dev).reshape(N, nd).clone()
        centroid = engine.as_device_tensor(mass_centroid, f32, dev).reshape(nd, N).clone()
        angle = engine.as_device_tensor(mass_angle, f32, dev).reshape(N).clone()
        plan = engine.Plan.get(world_size=self.world_size, nb_channels=C, slots=tuple(range(K)), c_in=(0, ) * K, gf_ids=(0, ) * K,
                               nb_slots=K, state_fn='v1', weighted_average=True, R=self.R, stats_dt=self.dt, device=dev)
        stats = torch.empty((_lib.LNX_NB_STATS, N), dtype=f32, device=dev)
        cm = torch.empty((N, C), dtype=f32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(plan.lib.lnx_compute_stats(plan.handle, N, cells.data_ptr(), field.data_ptr(), potential.data_ptr(), shift.data_ptr(),
                                                  centroid.data_ptr(), angle.data_ptr(), stats.data_ptr(), cm.data_ptr(), stream))
        out = {k: stats[i] for i, k in enumerate(_lib.STAT_KEYS)}
        out['channel_mass'] = cm
        return out, shift, centroid, angle


def build_compute_stats_fn(world_params: Dict, render_params: Dict) -> ComputeStatsFn:
    """statistics.py:11-33: R and dt = 1/T come from the build-time config (not from the per-solution T)."""
    return ComputeStatsFn(tuple(render_params['world_size']), float(world_params['R']), 1. / float(world_params['T']))


# ---- host-side heuristics on [T, N] statistics (used by runner.run's python-loop semantics and for tests) ----
def monotonic_heuristic(sign, previous_sign, monotone_counter):  # statistics.py:287-306
    monotone_counter = monotone_counter * (sign == previous_sign) + 1
    return monotone_counter <= MONOTONIC_STOP_STEP, monotone_counter


def mass_volume_heuristic(mass_volume, mass_volume_counter):  # statistics.py:317-333
    mass_volume_counter = mass_volume_counter * (mass_volume > MASS_VOLUME_THRESHOLD) + 1
    return mass_volume_counter <= MASS_VOLUME_STOP_STEP, mass_volume_counter


def min_mass_heuristic(epsilon, mass):  # statistics.py:254-266
    return mass >= epsilon


def max_mass_heuristic(init_mass, mass):  # statistics.py:269-281
    return mass <= 3 * init_mass


def check_heuristics(stats: Dict[str, torch.Tensor]) -> torch.Tensor:
    """statistics.py:134-205 on tensors ``[T, N]`` wherever they live (the scan kernels compute the same thing in-kernel and
    return its time-sum as ``stats['N']``; this version serves callers that post-process stored statistics and the
    direct-convolution scans)."""
    mass, cm, mv = stats['mass'].detach(), stats['channel_mass'].detach(), stats['mass_volume'].detach()
    T, N = mass.shape
    dev = mass.device
    should_continue = torch.ones(N, device=dev)
    init_cm, prev_mass, prev_sign = cm[0], mass[0], torch.zeros(N, device=dev)
    mono = torch.zeros(N, dtype=torch.int32, device=dev)
    vol = torch.zeros(N, dtype=torch.int32, device=dev)
    out = torch.empty((T, N), device=dev)
    for t in range(T):
        cond = (cm[t] >= EPSILON).all(dim=1) & (cm[t] <= 3 * init_cm).all(dim=1)
        sign = torch.sign(mass[t] - prev_mass)
        c, mono = monotonic_heuristic(sign, prev_sign, mono)
        cond = cond & c
        c, vol = mass_volume_heuristic(mv[t], vol)
        cond = cond & c
        should_continue = should_continue * cond
        prev_mass, prev_sign = mass[t], sign
        out[t] = should_continue
    return out
