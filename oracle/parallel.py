"""Test infrastructure: the oracle's batched scan on all host cores, with an early exit per world.

``lo.run_scan`` simulates every world for ``max_run_iter`` steps.  For the integer-parity checks only ``N`` (the index of
the first step whose stop criteria fail, runner.py:161-162 + statistics.py:134-205) and the rows ``[ns-128, ns)`` with
``ns = max(N, 128)`` (qd.py:181-185) matter, and worlds are independent, so a world is dropped from the oracle's batch once
its ``should_continue`` has fallen to 0 and it has at least 128 rows.  ``tests/test_host_logic.py`` checks that this gives
the same ``N`` and the same rows as ``lo.run_scan``.
"""
import multiprocessing as mp
import os
from typing import Dict, Sequence

import numpy as np

from . import lenia_oracle as lo


class _Online:
    """check_heuristics (statistics.py:134-205) one step at a time, vectorised over the worlds still in the batch."""

    def __init__(self, n, dtype):
        self.should_continue = np.ones(n, dtype)
        self.init_cm = None
        self.prev_mass = None
        self.prev_sign = np.zeros(n, dtype)
        self.mono = np.zeros(n, np.int32)
        self.vol = np.zeros(n, np.int32)
        self.n_alive = np.zeros(n, dtype)

    def step(self, st):
        mass, cm = st['mass'], st['channel_mass']
        eps = mass.dtype.type(lo.EPSILON)
        if self.init_cm is None:
            self.init_cm, self.prev_mass = cm, mass
        cond = (cm >= eps).all(axis=1) * (cm <= 3 * self.init_cm).all(axis=1)
        with np.errstate(invalid='ignore'):
            sign = np.sign(mass - self.prev_mass)
        c, self.mono = lo.monotonic_heuristic(sign, self.prev_sign, self.mono)
        cond = cond * c
        c, self.vol = lo.mass_volume_heuristic(st['mass_volume'], self.vol)
        cond = cond * c
        self.should_continue = self.should_continue * cond
        self.prev_mass, self.prev_sign = mass, sign
        self.n_alive = self.n_alive + self.should_continue

    def keep(self, sel):
        for k in ('should_continue', 'init_cm', 'prev_mass', 'prev_sign', 'mono', 'vol', 'n_alive'):
            setattr(self, k, getattr(self, k)[sel])


def scan_until_decided(cells0, K, gf_params, weights, T, max_run_iter, update_fn, compute_stats_fn, keys: Sequence[str],
                       window: int = lo.NB_STATS_STEPS) -> Dict[str, np.ndarray]:
    """One solution, ``cells0 [n, C, *dims]``.  Returns ``N [n]``, ``steps [n]`` (steps actually simulated) and, per key, the mean
    of rows ``[ns - window, ns)``, ``ns = max(N, window)`` clamped to the rows that exist (qd.py:181-185)."""
    dtype = cells0.dtype
    n = cells0.shape[0]
    dt = dtype.type(1.) / dtype.type(T)
    shift, centroid, angle = lo._init_carry(cells0, dtype)
    online = _Online(n, dtype)
    alive_idx = np.arange(n)
    rows = {k: np.zeros((max_run_iter, n), dtype) for k in keys}
    out_N, out_steps = np.zeros(n, dtype), np.zeros(n, np.int32)
    cells = cells0
    for t in range(max_run_iter):
        new_cells, field, potential = update_fn(cells, K, gf_params, weights, dt)
        st, shift, centroid, angle = compute_stats_fn(cells, field, potential, shift, centroid, angle)
        online.step(st)
        for k in keys:
            rows[k][t, alive_idx] = st[k]
        cells = new_cells
        done = (online.should_continue == 0) & (t + 1 >= min(window, max_run_iter))
        if t + 1 == max_run_iter:
            done[:] = True
        if done.any():
            out_N[alive_idx[done]] = online.n_alive[done]
            out_steps[alive_idx[done]] = t + 1
            sel = ~done
            if not sel.any():
                break
            alive_idx = alive_idx[sel]
            cells, shift, centroid, angle = cells[sel], shift[sel], centroid[:, sel], angle[sel]
            online.keep(sel)
    res = {'N': out_N, 'steps': out_steps}
    w = min(window, max_run_iter)
    ns = np.clip(out_N.astype(np.int64), w, max_run_iter)
    for k in keys:
        res[k] = np.array([rows[k][max(ns[i] - window, 0):ns[i], i].mean(dtype=np.float64) for i in range(n)], dtype=dtype)
    return res


def _worker(job):
    (kernels_params, world_size, nb_channels, R, T, state_fn, average, wp, rp, cells0, steps, keys, dt_name, gf_params, weights) = job
    dtype = np.float64 if dt_name == 'f64' else np.float32
    K, mapping = lo.get_kernels_and_mapping(kernels_params, world_size, nb_channels, R, True, dtype)
    update_fn = lo.build_update_fn(mapping, state_fn, average)
    stats_fn = lo.build_compute_stats_fn(wp, rp, dtype)
    gfp = mapping.get_gf_params(dtype) if gf_params is None else np.asarray(gf_params, dtype)
    w = mapping.get_kernels_weight_per_channel(dtype) if weights is None else np.asarray(weights, dtype)
    return scan_until_decided(cells0.astype(dtype), K, gfp, w, dtype(T), steps, update_fn, stats_fn, keys)


def parallel_scan(kernels_params, world_params, render_params, cells0, steps, keys, dtype='f32', chunk=8, procs=None,
                  gf_params=None, weights=None, T=None):
    """All worlds of ONE solution over a process pool (chunks of ``chunk`` worlds).  ``cells0 [n, C, *dims]`` float32."""
    import copy
    wp, rp = dict(world_params), dict(render_params)
    n = cells0.shape[0]
    jobs = []
    for a in range(0, n, chunk):
        jobs.append((copy.deepcopy(kernels_params), list(rp['world_size']), wp['nb_channels'], wp['R'], wp['T'] if T is None else T,
                     wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), wp, rp, cells0[a:a + chunk], steps, tuple(keys),
                     dtype, gf_params, weights))
    procs = procs or min(len(jobs), os.cpu_count() or 1)
    if procs <= 1:
        parts = [_worker(j) for j in jobs]
    else:
        with mp.get_context('spawn').Pool(procs) as pool:
            parts = pool.map(_worker, jobs, chunksize=1)
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
