#!/usr/bin/env python
"""Secondary measurements: the BASELINE configs other than the headline one (A, the search form of B, C, D, E).  One JSON line per config.

    python tools/bench_configs.py [--configs A,B,C,D,E] [--steps N]

These are parity-test configurations first (tests/test_gpu_parity.py); this script only times them (CUDA events around
the whole run_scan* call, inputs resident in HBM) and relates the result to the roofline SURVEY.md §8d assigns.
"""
import argparse
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leniax_b200 import _lib, helpers, initializations, kernels, lenia, loader, runner, statistics, utils  # noqa: E402

DEV = 'cuda:0'
FP32_PEAK = None


def timed(fn, reps=3):
    # two untimed calls with the previous result still alive, exactly like the timed loop below: torch's caching allocator then
    # owns both generations of output / workspace blocks and no cudaMalloc (an implicit device synchronisation, 5-40 ms measured on
    # the 10 ms call of config D, tools/time_config_d_calls.py) falls inside the timed region
    out = fn()
    out = fn()  # noqa: F841
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def fp32_peak():
    global FP32_PEAK
    if FP32_PEAK is None:
        lib = _lib.load_library()
        tf, ms = _lib.ctypes.c_double(), _lib.ctypes.c_double()
        _lib.check(lib.lnx_measure_fp32_peak(4096, _lib.ctypes.byref(tf), _lib.ctypes.byref(ms), None))
        FP32_PEAK = tf.value
    return FP32_PEAK


def orbium_parts(size=128, R=13):
    kp = [dict(k_slug='circle_2d', k_params=[1., [1.]], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015],
               h=1., c_in=0, c_out=0)]
    K, mapping = kernels.get_kernels_and_mapping(kp, [size, size], 1, R, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': R, 'T': 10}, {'world_size': [size, size]})
    return K, mapping, ufn, sfn


def orbium_cells():
    cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', 'orbium.yaml'))
    return loader.load_raw_cells(cfg, use_init_cells=False).numpy()[0]


def config_a(steps):
    """Orbium 1c1k 128x128, one world, python-loop semantics (runner.run), 1024 steps."""
    K, mapping, ufn, sfn = orbium_parts()
    world = np.zeros((1, 1, 128, 128), np.float32)
    world[0, 0, 54:74, 54:74] = orbium_cells()
    cells = torch.from_numpy(world).to(DEV)
    gf, w = mapping.get_gf_params(DEV), mapping.get_kernels_weight_per_channel(DEV)
    ms, out = timed(lambda: runner.run(None, cells, K, gf, w, 10., steps, 13, ufn, sfn, stat_trunc=True))
    ms2, _ = timed(lambda: runner.run_scan_mem_optimized(None, cells[None], K[None], gf[None], w[None], torch.tensor([10.], device=DEV), steps, 13, ufn, sfn))
    cu = 128 * 128 * steps
    return {'config': 'A: Orbium 1c1k 128x128 single world, runner.run semantics (full trajectory kept)', 'steps': steps, 'ms': ms,
            'cell_updates_per_s': cu / (ms * 1e-3), 'N': int(out[3]['N']),
            'stats_only_ms': ms2, 'stats_only_cell_updates_per_s': cu / (ms2 * 1e-3),
            'note': 'one world occupies one SM (one CTA per world): latency, not throughput, configuration'}


def config_b(steps):
    """BASELINE configs[1] as search_for_init_mem_optimized runs it: one solution x 4096 perlin
    initialisations, all statistics, with and without the early-stop extension.  Most perlin soups die early, so the
    early-stop time is what a search pays; the headline bench (bench.py) uses all-surviving worlds instead."""
    n_init = 4096
    K, mapping, ufn, sfn = orbium_parts()
    _, noise = initializations.perlin(initializations.RngKey(1), n_init, [128, 128], 13, [.15, .015], device=DEV)
    cells = noise.reshape(1, n_init, 1, 128, 128)
    args = (cells, K[None], mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None], torch.full((1, ), 10., device=DEV))
    ms, out = timed(lambda: runner.run_scan_mem_optimized(None, *args, steps, 13, ufn, sfn))
    ms_early, out_e = timed(lambda: runner.run_scan_mem_optimized(None, *args, steps, 13, ufn, sfn, early_stop=True))
    n_full, n_early = out[0]['N'], out_e[0]['N']
    cu = n_init * 128 * 128 * steps
    tfl = cu * 110 / (ms * 1e-3) / 1e12
    return {'config': 'B (search form): 1 solution x 4096 perlin inits, Orbium physics 1c1k 128x128, all statistics', 'steps': steps, 'ms': ms,
            'cell_updates_per_s': cu / (ms * 1e-3), 'early_stop_ms': ms_early,
            'early_stop_note': 'extension (early_stop=True): a world stops once its stop criteria fired and 128 rows exist; N is unchanged',
            'N_identical_with_early_stop': bool(torch.equal(n_full, n_early)), 'mean_N': float(n_full.mean()), 'max_N': float(n_full.max()),
            'survivors': int((n_full >= steps).sum()),
            'roofline': {'bound': 'fp32', 'flop_per_cell_update': 110, 'achieved_tflops': tfl, 'peak_tflops': fp32_peak(), 'frac': tfl / fp32_peak()}}


def config_c(steps):
    """3c6k (conf/config_qd_cmame_3c6k.yaml physics), 16 solutions x 128 inits = 2048 worlds, per-solution parameters."""
    pairs = [(0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 0)]
    bs = {(0, 0): [1.], (1, 1): [.5, 1.], (2, 2): [1., .5]}
    base = [dict(k_slug='circle_2d', k_params=[1., bs.get(p, [1.])], kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4',
                 gf_params=[.17, .015], h=1., c_in=p[0], c_out=p[1]) for p in pairs]
    rng = np.random.default_rng(2)
    n_sols, n_init = 16, 128
    Ks, gfs, ws, cells = [], [], [], []
    key = initializations.RngKey(2)
    for s in range(n_sols):
        kp = copy.deepcopy(base)
        for k in kp:
            g = rng.random(3)
            k['gf_params'] = [round(.1 + .4 * g[0], 8), round(.005 + .095 * g[1], 8)]  # s kept > 0 for the timing run
            k['h'] = round(.05 + .95 * g[2], 8)
        K, mapping = kernels.get_kernels_and_mapping(kp, [128, 128], 3, 13, device=DEV)
        Ks.append(K)
        gfs.append(mapping.get_gf_params(DEV))
        ws.append(mapping.get_kernels_weight_per_channel(DEV))
        key, noise = initializations.perlin(key, 3 * n_init, [128, 128], 13, kp[0]['gf_params'], device=DEV)
        cells.append(noise.reshape(n_init, 3, 128, 128))
    ufn = helpers.build_update_fn(Ks[0].shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': 13, 'T': 10}, {'world_size': [128, 128]})
    args = (torch.stack(cells), torch.stack(Ks), torch.stack(gfs), torch.stack(ws), torch.full((n_sols, ), 10., device=DEV))
    ms, out = timed(lambda: runner.run_scan_mem_optimized(None, *args, steps, 13, ufn, sfn), reps=2)
    ms_early, _ = timed(lambda: runner.run_scan_mem_optimized(None, *args, steps, 13, ufn, sfn, early_stop=True), reps=2)
    cu = n_sols * n_init * 128 * 128 * steps
    tfl = cu * 486 / (ms * 1e-3) / 1e12
    return {'config': 'C: 3 channels 6 kernels 128x128, 16 solutions x 128 perlin inits, generic resident kernel', 'steps': steps, 'ms': ms,
            'early_stop_ms': ms_early, 'early_stop_note': 'extension (early_stop=True): a world stops once its stop criteria fired and 128 rows exist; '
            'everything qd.update_individuals reads is unchanged.  Not used for the throughput figure (it skips steps).',
            'cell_updates_per_s': cu / (ms * 1e-3), 'roofline': {'bound': 'fp32', 'flop_per_cell_update': 486, 'achieved_tflops': tfl,
                                                               'peak_tflops': fp32_peak(), 'frac': tfl / fp32_peak()},
            'mean_N': float(out[0]['N'].mean())}


def config_d(steps):
    """One 2048x2048 world, R=52 (kernel 104x104), Orbium upscaled x4 at 16 positions; tiled engine, HBM-bound roofline."""
    size, scale, R = 2048, 4, 52
    K, mapping, ufn, sfn = orbium_parts(size, R)
    big = np.kron(orbium_cells(), np.ones((scale, scale), np.float32))
    world = np.zeros((size, size), np.float32)
    rng = np.random.default_rng(4)
    for _ in range(16):
        y, x = rng.integers(0, size - big.shape[0], 2)
        world[y:y + big.shape[0], x:x + big.shape[1]] = np.maximum(world[y:y + big.shape[0], x:x + big.shape[1]], big)
    cells = torch.from_numpy(world).to(DEV)[None, None, None]
    gf, w = mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None]
    ms, out = timed(lambda: runner.run_scan_mem_optimized(None, cells, K[None], gf, w, torch.tensor([10.], device=DEV), steps, R, ufn, sfn), reps=5)
    cu = size * size * steps
    gbs = cu * 32 / (ms * 1e-3) / 1e9
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    hbm = peaks.get('hbm_gbs', 6650.0)
    engine = ('generic tiled passes' if runner.TILED_GENERIC else 'four-step warp-per-line passes (lnx_tiled2k.cuh), '
              + ('real row per warp' if runner.T2K_REAL_ROWS else 'row pair per warp') + ', CUDA-graph step loop')
    return {'config': 'D: one 2048x2048 world, 1c1k, R=52, ' + engine, 'steps': steps, 'ms': ms, 'cell_updates_per_s': cu / (ms * 1e-3),
            'roofline': {'bound': 'hbm', 'bytes_per_cell_update': 32, 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                         'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'},
            'N': float(out[0]['N'][0, 0]), 'note': 'the 48 MiB working set fits the 126 MB L2: algorithmic bytes, not DRAM traffic'}


def config_e(steps):
    """256 worlds 64^3, 1c1k, R=13 raw spherical-shell kernel, uniform random init."""
    D, R, n = 64, 13, 256
    kern = kernels.sphere_nd(R, [1., [1.]], 'poly_quad', [4], device=DEV)
    kp = [dict(k_slug='raw', k_params=kern, kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015], h=1., c_in=0, c_out=0)]
    K, mapping = kernels.get_kernels_and_mapping(kp, [D, D, D], 1, R, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': R, 'T': 10}, {'world_size': [D, D, D]})
    _, cells = initializations.random_uniform(initializations.RngKey(5), n, [D, D, D], R, [.15, .015], device=DEV)
    cells = cells[None, :, None]
    gf, w = mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None]
    ms, out = timed(lambda: runner.run_scan_mem_optimized(None, cells, K[None], gf, w, torch.tensor([10.], device=DEV), steps, R, ufn, sfn), reps=int(os.environ.get('LNX_BENCH_REPS', 2)))
    cu = n * D**3 * steps
    tfl = cu * 136 / (ms * 1e-3) / 1e12
    gbs = cu * 32 / (ms * 1e-3) / 1e9
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    hbm = peaks.get('hbm_gbs', 6650.0)
    engine = ('generic tiled passes' if runner.TILED_GENERIC else 'thread-per-line passes (lnx_tiled64.cuh)' if runner.T64_LINE
              else 'half-line passes (lnx_tiled64h.cuh), ' + ('persistent whole-scan kernel' if runner.T64_WHOLE_SCAN else 'one launch per pass and step'))
    return {'config': 'E: 256 worlds 64^3, 1c1k, ' + engine, 'steps': steps, 'ms': ms, 'cell_updates_per_s': cu / (ms * 1e-3),
            'roofline': {'bound': 'fp32 (SURVEY) / hbm (multi-pass engines: 256 MB of state > L2)', 'flop_per_cell_update': 136, 'achieved_tflops': tfl,
                         'peak_tflops': fp32_peak(), 'frac': tfl / fp32_peak(), 'achieved_gbs_at_32B': gbs, 'hbm_peak_gbs': hbm,
                         'frac_hbm': gbs / hbm}, 'mean_N': float(out[0]['N'].mean())}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='A,C,D,E')
    ap.add_argument('--steps', type=int, default=0)
    ap.add_argument('--tiled-generic', action='store_true', help='configs D / E through the generic tiled passes (A/B run)')
    ap.add_argument('--t2k-real-rows', action='store_true', help='config D through the rows kernels with one real row per warp (A/B run)')
    ap.add_argument('--t64-whole-scan', action='store_true', help='config E through the persistent whole-scan kernel (default only up to 128 worlds; A/B run)')
    ap.add_argument('--t64-line', action='store_true', help='config E through the round-1 thread-per-line step kernels (A/B run)')
    a = ap.parse_args()
    runner.TILED_GENERIC = a.tiled_generic
    runner.T64_LINE = a.t64_line
    runner.T64_WHOLE_SCAN = a.t64_whole_scan
    runner.T2K_REAL_ROWS = a.t2k_real_rows
    default_steps = {'A': 1024, 'B': 1024, 'C': 1024, 'D': 256, 'E': 64}
    fns = {'A': config_a, 'B': config_b, 'C': config_c, 'D': config_d, 'E': config_e}
    for c in a.configs.split(','):
        print(json.dumps(fns[c](a.steps or default_steps[c])), flush=True)
