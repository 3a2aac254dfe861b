"""Time loops (reference: leniax/runner.py:16-334), executed by the persistent CUDA kernels.

Same names and positional signatures as the reference.  ``update_fn`` is a ``core.UpdateFn`` descriptor and
``compute_stats_fn`` a ``statistics.ComputeStatsFn`` descriptor (both built by ``helpers`` / ``statistics`` exactly like
the reference builds its callables).  Arrays may be torch tensors (CUDA or CPU) or numpy arrays; results are CUDA
tensors.  ``rng_key`` is accepted and ignored: no registered state function consumes randomness (core.py:245-319).
"""
import weakref
from typing import Dict, Tuple

import numpy as np

import torch

from . import _lib, engine
from .constant import EPSILON, START_CHECK_STOP
from .core import UpdateFn
from .statistics import (MASS_VOLUME_STOP_STEP, MASS_VOLUME_THRESHOLD, MONOTONIC_STOP_STEP, ComputeStatsFn)


def _check_fns(update_fn, compute_stats_fn):
    if not isinstance(update_fn, UpdateFn):
        raise NotImplementedError(
            'update_fn must be the descriptor returned by leniax_b200.helpers.build_update_fn; arbitrary Python '
            'callables cannot be fused into the CUDA scan and there is no CPU fallback'
        )
    if not isinstance(compute_stats_fn, ComputeStatsFn):
        raise NotImplementedError('compute_stats_fn must come from leniax_b200.statistics.build_compute_stats_fn')


_PARAM_SUMMARY_CACHE: Dict[int, tuple] = {}  # id(base tensor of the weights) -> (weakref to it, key, result)


def _base_of(t: torch.Tensor) -> torch.Tensor:
    return t._base if t._base is not None else t


def _param_summary(gf_params: torch.Tensor, weights: torch.Tensor, average: bool) -> Tuple[bool, Tuple[int, ...]]:
    """What the launch needs to know about the device-side parameters, from ONE device reduction and ONE host sync (cached per
    tensor OBJECT and version — views are looked up through their base — so a loop over the same parameters syncs once):

    * ``finite``: NaN cannot be born in the step — no growth width ``s == 0``, no all-zero weight row, nothing non-finite;
    * ``c_out[k]``: the single channel whose weight is non-zero in column ``k`` for some solution (``kernels.py:110-111`` writes
      ``W[c_out][k] = h``), ``LNX_COUT_NONE`` for an all-zero column; ``LNX_COUT_ANY`` everywhere when some column feeds several
      channels (hand-made weights): the engine then uses the weights tensor as given.

    ``gf_params [S, K, 2]``, ``weights [S, C, K]``."""
    wb, gb = _base_of(weights), _base_of(gf_params)
    key = (wb._version, tuple(weights.shape), weights.storage_offset(), id(gb), gb._version, tuple(gf_params.shape),
           gf_params.storage_offset(), bool(average))
    hit = _PARAM_SUMMARY_CACHE.get(id(wb))
    if hit is not None and hit[0]() is wb and hit[1]() is gb and hit[2] == key:
        return hit[3]
    C, K = weights.shape[-2], weights.shape[-1]
    bad = (~torch.isfinite(gf_params).all()) | (gf_params[..., 1] == 0).any() | (~torch.isfinite(weights).all())
    if average:
        bad = bad | (weights.sum(dim=-1) == 0).any()
    nz = (weights.reshape(-1, C, K) != 0).any(dim=0)  # [C, K]
    packed = torch.cat([bad.reshape(1), nz.reshape(-1)]).to(torch.uint8).cpu().tolist()  # the one synchronisation
    finite = packed[0] == 0
    c_out = []
    for k in range(K):
        rows = [c for c in range(C) if packed[1 + c * K + k]]
        c_out.append(rows[0] if len(rows) == 1 else (_lib.LNX_COUT_NONE if not rows else None))
    res = (finite, tuple(_lib.LNX_COUT_ANY for _ in range(K)) if any(v is None for v in c_out) else tuple(c_out))
    wid = id(wb)
    _PARAM_SUMMARY_CACHE[wid] = (weakref.ref(wb, lambda _r, wid=wid: _PARAM_SUMMARY_CACHE.pop(wid, None)), weakref.ref(gb), key, res)
    return res


GENERIC_OLD = False  # tests: several channels / kernels through the older lnx_world128_generic kernel (cross-check)
GENERIC_1CTA = False  # tests / A-B runs: several channels / kernels through lnx_world128_gen_tm (one world per SM) instead of gen2
TILED_GENERIC = False  # tests / A-B runs: 64^3 one-channel one-kernel worlds through the generic tiled passes instead of lnx_tiled64.cuh
T64_LINE = False  # tests / A-B runs: 64^3 worlds through the round-1 thread-per-line step kernels instead of the half-line kernels (lnx_tiled64h.cuh)
T2K_REAL_ROWS = False  # tests / A-B runs: 2048^2 worlds through the rows kernels with one real row per warp
T64_STEPWISE = False  # tests / A-B runs: 64^3 worlds with one launch per pass and step whatever the number of worlds (default above 128)
T64_WHOLE_SCAN = False  # ... with the persistent whole-scan kernel whatever the number of worlds (default up to 128)
FORCE_TILED_ENGINE = False  # tests set this to run 128x128 worlds through the tiled multi-pass engine as a cross-check


def _scan(cells0, K, gf_params, weights, T, max_run_iter, update_fn: UpdateFn, stats_fn: ComputeStatsFn, *, batched: bool,
          keep_trajectory: bool, early_stop: bool = False, want_final_cells: bool = True):
    assert max_run_iter > 0, f"max_run_iter must be positive, value given: {max_run_iter}"  # runner.py:51
    dev = engine.require_cuda_device(cells0.device if isinstance(cells0, torch.Tensor) and cells0.is_cuda else None)
    f32 = torch.float32
    host_cells = cells0 if (isinstance(cells0, torch.Tensor) and not cells0.is_cuda and cells0.dtype == f32 and batched
                            and cells0.is_contiguous()) else None
    if host_cells is None:
        cells0 = engine.as_device_tensor(cells0, f32, dev)
    gf_params = engine.as_device_tensor(gf_params, f32, dev)
    weights = engine.as_device_tensor(weights, f32, dev)
    T = engine.as_device_tensor(T, f32, dev)
    K = engine.as_device_tensor(K, torch.complex64 if update_fn.get_potential_fn.fft else f32, dev)
    if not batched:  # (host_cells is only kept for batched calls)
        cells0, gf_params, weights, T, K = cells0[None], gf_params[None], weights[None], T.reshape(1), K[None]
    shape = tuple(host_cells.shape) if host_cells is not None else tuple(cells0.shape)
    n_sols, n_init, C = shape[0], shape[1], shape[2]
    world_size = shape[3:]
    if tuple(stats_fn.world_size) != world_size:
        raise ValueError(f'compute_stats_fn was built for world_size {stats_fn.world_size}, cells are {world_size}')
    pf = update_fn.get_potential_fn
    if not pf.fft:
        if host_cells is not None:
            cells0 = host_cells.to(dev)
        return _scan_conv(cells0, K, gf_params, weights, T, max_run_iter, update_fn, stats_fn, keep_trajectory)
    from . import kernels as _kernels
    if not _kernels.is_pow2_world(world_size):
        # The FFT engines need powers of two; the reference's fftn takes any size (core.py:81).  Such worlds (2-D) are stepped by direct
        # convolution with the taps recovered from K, statistics by lnx_compute_stats: a host loop like the fft=False path - a
        # compatibility path, not a throughput path.
        import dataclasses
        if host_cells is not None:
            cells0 = host_cells.to(dev)
        taps = [_kernels.spatial_from_spectrum(K[s], pf.nb_slots, world_size) for s in range(n_sols)]
        kh, kw = max(t.shape[2] for t in taps), max(t.shape[3] for t in taps)
        taps = [torch.nn.functional.pad(t, ((kw - t.shape[3]) // 2, ) * 2 + ((kh - t.shape[2]) // 2, ) * 2) for t in taps]  # odd sizes: centred
        conv_fn = UpdateFn(dataclasses.replace(pf, fft=False), update_fn.get_field_fn, update_fn.get_state_fn)
        return _scan_conv(cells0, torch.stack(taps), gf_params, weights, T, max_run_iter, conv_fn, stats_fn, keep_trajectory)
    slots, c_in, gf_ids = update_fn.kernel_layout(C)
    K = K.reshape((n_sols, pf.nb_slots) + world_size)
    gfp, wts = gf_params.reshape(n_sols, len(slots), 2), weights.reshape(n_sols, C, len(slots))
    finite, c_out = _param_summary(gfp, wts, update_fn.get_field_fn.average)
    plan = engine.Plan.get(world_size=world_size, nb_channels=C, slots=slots, c_in=c_in, gf_ids=gf_ids, nb_slots=pf.nb_slots,
                           state_fn=update_fn.get_state_fn.slug, weighted_average=update_fn.get_field_fn.average, R=stats_fn.R,
                           stats_dt=stats_fn.dt, device=dev, force_tiled=FORCE_TILED_ENGINE, c_out=c_out)
    flags = 0
    if early_stop:
        flags |= _lib.LNX_RUN_EARLY_STOP
    if GENERIC_OLD:
        flags |= _lib.LNX_RUN_GENERIC_OLD
    if GENERIC_1CTA:
        flags |= _lib.LNX_RUN_GENERIC_1CTA
    if TILED_GENERIC:
        flags |= _lib.LNX_RUN_TILED_GENERIC
    if T64_LINE:
        flags |= _lib.LNX_RUN_T64_LINE
    if T2K_REAL_ROWS:
        flags |= _lib.LNX_RUN_T2K_REAL_ROWS
    if T64_STEPWISE:
        flags |= _lib.LNX_RUN_T64_STEPWISE
    if T64_WHOLE_SCAN:
        flags |= _lib.LNX_RUN_T64_WHOLE_SCAN
    if finite:
        flags |= _lib.LNX_RUN_ASSUME_FINITE
    if c_out[0] != _lib.LNX_COUT_ANY:
        flags |= _lib.LNX_RUN_WEIGHTS_MATCH_COUT
    dt = (1. / T.reshape(n_sols)).contiguous()  # runner.py:307
    if host_cells is not None:
        n_first = 2 * torch.cuda.get_device_properties(dev).multi_processor_count  # one full wave of CTAs of the fused kernel
        if n_sols == 1 and not keep_trajectory and n_init >= 4 * n_first and host_cells.is_pinned():  # (pageable copies block the host: no overlap)
            return _scan_pipelined_upload(plan, host_cells, n_first, K, gfp, wts, dt, max_run_iter, flags, dev, want_final_cells)
        cells0 = host_cells.to(dev, non_blocking=True)
    return plan.run_scan(cells0.contiguous(), K, gfp, wts, dt, max_run_iter, keep_trajectory=keep_trajectory, flags=flags,
                         want_final_cells=want_final_cells)


_COPY_STREAMS: Dict[str, torch.cuda.Stream] = {}


def _scan_pipelined_upload(plan, host_cells, n_first, K, gfp, wts, dt, max_run_iter, flags, dev, want_final_cells=True):
    """Initial states given in PINNED host memory (one solution, many initialisations): the first wave of worlds is uploaded and
    started at once, the rest of the batch is uploaded on a copy stream while that wave computes, then runs as a second
    launch.  Worlds are independent, so the two launches give bit-identical rows to a single one."""
    main = torch.cuda.current_stream(dev)
    side = _COPY_STREAMS.setdefault(str(dev), torch.cuda.Stream(device=dev))
    with torch.cuda.device(dev):
        side.wait_stream(main)
        first = host_cells[:, :n_first].to(dev, non_blocking=True)
        with torch.cuda.stream(side):
            rest = host_cells[:, n_first:].to(dev, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(side)
        r1 = plan.run_scan(first, K, gfp, wts, dt, max_run_iter, keep_trajectory=False, flags=flags, want_final_cells=want_final_cells)
        main.wait_event(uploaded)
        rest.record_stream(main)
        r2 = plan.run_scan(rest, K, gfp, wts, dt, max_run_iter, keep_trajectory=False, flags=flags, want_final_cells=want_final_cells)
    stats = {}
    for k, v in r1['stats'].items():
        axis = 2 if k == 'channel_mass' else (1 if k == 'N' else 2)  # [S, T, I, C] / [S, I] / [S, T, I]
        stats[k] = torch.cat([v, r2['stats'][k]], dim=axis)
    final = torch.cat([r1['final_cells'], r2['final_cells']], dim=1) if want_final_cells else None
    return {'stats': stats, 'final_cells': final, 'cells': None, 'field': None,
            'potential': None}


def _scan_conv(cells0, K, gf_params, weights, T, max_run_iter, update_fn: UpdateFn, stats_fn: ComputeStatsFn, keep_trajectory: bool):
    """The scan with the direct-convolution potential (``fft=False``): the reference's cross-check path, run as a host loop
    of ``lnx_update_conv`` + ``lnx_compute_stats`` per step (runner._scan_fn, runner.py:295-334: statistics of the
    PRE-update cells with this step's field / potential), ``check_heuristics`` afterwards (runner.py:161-162)."""
    from . import core
    from .statistics import check_heuristics
    n_sols, n_init, C = cells0.shape[:3]
    dev = cells0.device
    nd = cells0.dim() - 3
    per_sol = []
    traj = {'cells': [], 'field': [], 'potential': []}
    finals = []
    for s in range(n_sols):
        cells = cells0[s].contiguous()
        shift = torch.zeros((n_init, nd), dtype=torch.int32, device=dev)  # runner.py:271-292
        centroid = torch.zeros((nd, n_init), dtype=torch.float32, device=dev)
        angle = torch.zeros((n_init, ), dtype=torch.float32, device=dev)
        rows, tc, tf, tp = [], [], [], []
        for _ in range(max_run_iter):
            new_cells, field, potential = core.update_conv(cells, K[s], gf_params[s], weights[s], 1. / T.reshape(-1)[s], update_fn)
            st, shift, centroid, angle = stats_fn(cells, field, potential, shift, centroid, angle)
            rows.append(st)
            if keep_trajectory:
                tc.append(cells)
                tf.append(field)
                tp.append(potential)
            cells = new_cells
        stats = {k: torch.stack([r[k] for r in rows]) for k in rows[0]}  # [T, N] / [T, N, C]
        stats['N'] = check_heuristics(stats).sum(dim=0)
        per_sol.append(stats)
        finals.append(cells)
        if keep_trajectory:
            traj['cells'].append(torch.stack(tc))
            traj['field'].append(torch.stack(tf))
            traj['potential'].append(torch.stack(tp))
    out = {k: torch.stack([p[k] for p in per_sol]) for k in per_sol[0]}
    res = {'stats': out, 'final_cells': torch.stack(finals), 'cells': None, 'field': None, 'potential': None}
    if keep_trajectory:
        for k in traj:
            res[k] = torch.stack(traj[k])
    return res


def run_scan(rng_key, cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float, update_fn,
             compute_stats_fn) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
    """Simulate a single configuration with ``N_init`` initialisations (runner.py:119-164).

    Returns ``(cells [T, N_init, C, H, W], field, potential [T, N_init, K, H, W], stats {k: [T, N_init]})`` with
    ``stats['channel_mass'] [T, N_init, C]`` and ``stats['N'] [N_init]``.
    """
    _check_fns(update_fn, compute_stats_fn)
    res = _scan(cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter, update_fn, compute_stats_fn, batched=False,
                keep_trajectory=True)
    stats = {k: v[0] for k, v in res['stats'].items()}
    return res['cells'][0], res['field'][0], res['potential'][0], stats


def run_scan_mem_optimized(rng_key, cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float,
                           update_fn, compute_stats_fn, early_stop: bool = False, return_final_cells: bool = True
                           ) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """Simulate ``N_sols`` configurations x ``N_init`` initialisations (runner.py:167-215).

    ``cells0 [N_sols, N_init, C, H, W]``, ``K [N_sols, 1, C, max_k, H, W]``, ``gf_params [N_sols, K, 2]``,
    ``kernels_weight_per_channel [N_sols, C, K]``, ``T [N_sols]``.  Returns ``(stats {k: [N_sols, T, N_init]}, final_cells)``.
    ``early_stop=True`` (extension) lets a world stop once its stop criteria fired and 128 rows exist; the rows the QD
    consumer reads (qd.py:181-185) are unaffected, later rows are zero.  ``return_final_cells=False`` (extension; the QD evaluation,
    which drops them, qd.py:63-66): the final states are neither allocated nor written, ``None`` is returned in their place.
    """
    _check_fns(update_fn, compute_stats_fn)
    res = _scan(cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter, update_fn, compute_stats_fn, batched=True,
                keep_trajectory=False, early_stop=early_stop, want_final_cells=return_final_cells)
    return res['stats'], res['final_cells']


def run_scan_mem_optimized_pmap(rng_key, cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float,
                                update_fn, compute_stats_fn) -> Tuple[Dict[str, torch.Tensor], torch.Tensor]:
    """Leading ``N_device`` axis (runner.py:218-268).  One process drives one GPU here, so the device axis is folded
    into the solution axis; multi-GPU sharding lives in ``leniax_b200.distributed``."""
    nd, ns = cells0.shape[0], cells0.shape[1]
    fold = lambda x: x.reshape((nd * ns, ) + tuple(x.shape[2:]))  # noqa: E731
    stats, final = run_scan_mem_optimized(rng_key, fold(cells0), fold(K), fold(gf_params), fold(kernels_weight_per_channel),
                                          fold(T), max_run_iter, R, update_fn, compute_stats_fn)
    unfold = lambda x: x.reshape((nd, ns) + tuple(x.shape[1:]))  # noqa: E731
    return {k: unfold(v) for k, v in stats.items()}, unfold(final)


def _heuristic_counter(keep: np.ndarray) -> np.ndarray:
    """All values of the recurrence ``c_t = c_{t-1} * keep_t + 1`` with ``c_{-1} = 0`` (statistics.py:287-306, 317-333)."""
    steps = np.arange(keep.shape[0])
    last_reset = np.maximum.accumulate(np.where(keep, 0, steps))
    return steps - last_reset + 1


def run(rng_key, cells, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float, update_fn, compute_stats_fn,
        stat_trunc: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
    """Simulate a single configuration with the semantics of the reference's python loop (runner.py:16-116).

    The reference dispatches two jitted calls per step and breaks on its on-the-fly heuristics.  The break only
    truncates, so the same result is obtained by scanning ``max_run_iter`` steps on the GPU and replaying the loop's
    stop rules (total-mass min/max against the un-normalised initial sum, monotone / volume counters that never see
    ``previous_mass`` updated, grace period ``START_CHECK_STOP``) on the statistics afterwards.
    """
    assert max_run_iter > 0, f"max_run_iter must be positive, value given: {max_run_iter}"
    assert cells.shape[0] == 1
    all_cells, all_fields, all_potentials, stats = run_scan(rng_key, cells, K, gf_params, kernels_weight_per_channel, T,
                                                           max_run_iter, R, update_fn, compute_stats_fn)
    stats.pop('N')
    mass = stats['mass'][:, 0].detach().cpu().numpy()
    mass_volume = stats['mass_volume'][:, 0].detach().cpu().numpy()
    init_mass = np.float32(all_cells[0].sum().item())  # runner.py:61 (not divided by R^2)
    # The loop's rules, vectorised over the steps (a Python loop over 1024 steps cost four times the simulation itself).
    # previous_mass and previous_sign are never updated in the reference loop (runner.py:62-63, 93-95): the sign is taken
    # against the initial sum and compared with 0.
    steps = np.arange(max_run_iter)
    counter = _heuristic_counter
    cond = (mass >= np.float32(EPSILON)) & (mass <= np.float32(3) * init_mass)
    cond &= counter(np.sign(mass - init_mass) == 0) <= MONOTONIC_STOP_STEP
    cond &= counter(mass_volume > np.float32(MASS_VOLUME_THRESHOLD)) <= MASS_VOLUME_STOP_STEP
    should_continue = np.logical_and.accumulate(cond)
    current_iter = max_run_iter - 1
    if stat_trunc is True:
        stopped = np.nonzero(~should_continue & (steps >= START_CHECK_STOP))[0]
        if stopped.size:
            current_iter = int(stopped[0])
    n = current_iter + 1
    stats = {k: v[:n] for k, v in stats.items()}
    stats['N'] = torch.tensor(current_iter)
    return all_cells[:n], all_fields[:n], all_potentials[:n], stats
