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
        raise NotImplementedError(
            'per-step statistics are fused into the scan kernel; call leniax_b200.runner.run / run_scan / '
            'run_scan_mem_optimized, which return the same stats dictionary'
        )


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
    """statistics.py:134-205 on host tensors ``[T, N]`` (the scan kernels compute the same thing on device and return
    its time-sum as ``stats['N']``; this host version exists for callers that post-process stored statistics)."""
    mass = stats['mass'].detach().cpu()
    cm = stats['channel_mass'].detach().cpu()
    mv = stats['mass_volume'].detach().cpu()
    T, N = mass.shape
    should_continue = torch.ones(N)
    init_cm, prev_mass, prev_sign = cm[0], mass[0], torch.zeros(N)
    mono = torch.zeros(N, dtype=torch.int32)
    vol = torch.zeros(N, dtype=torch.int32)
    out = torch.empty((T, N))
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
