#!/usr/bin/env python
"""A/B timing of BASELINE config E (256 worlds 64^3 x 64 steps) that survives a noisy box: the variants alternate inside ONE process and the
minimum / median over many repetitions are reported (a shared box time-slices the GPU: single runs vary by 2-3x).

    python tools/ab_config_e.py [--reps 20]      (LNX_T64_WINDOW is read once per process: run it again for another window)
"""
import argparse
import json
import os
import statistics as pystat
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leniax_b200 import helpers, initializations, kernels, runner, statistics  # noqa: E402

DEV = 'cuda:0'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--worlds', type=int, default=256)
    ap.add_argument('--steps', type=int, default=64)
    a = ap.parse_args()
    D, R, n = 64, 13, a.worlds
    kern = kernels.sphere_nd(R, [1., [1.]], 'poly_quad', [4], device=DEV)
    kp = [dict(k_slug='raw', k_params=kern, kf_slug='poly_quad', kf_params=[4], gf_slug='poly_quad4', gf_params=[.15, .015], h=1., c_in=0, c_out=0)]
    K, mapping = kernels.get_kernels_and_mapping(kp, [D, D, D], 1, R, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': R, 'T': 10}, {'world_size': [D, D, D]})
    _, cells = initializations.random_uniform(initializations.RngKey(5), n, [D, D, D], R, [.15, .015], device=DEV)
    cells = cells[None, :, None]
    gf, w = mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None]
    T = torch.tensor([10.], device=DEV)

    def run():
        return runner.run_scan_mem_optimized(None, cells, K[None], gf, w, T, a.steps, R, ufn, sfn)

    variants = {'whole_scan': dict(T64_WHOLE_SCAN=True), 'stepwise': dict(T64_STEPWISE=True), 'line64_round1': dict(T64_LINE=True)}
    times = {k: [] for k in variants}
    keep = None
    for rep in range(a.reps + 2):
        for name, flags in variants.items():
            for k, v in flags.items():
                setattr(runner, k, v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run()
            e1.record()
            torch.cuda.synchronize()
            for k in flags:
                setattr(runner, k, False)
            if rep >= 2:
                times[name].append(e0.elapsed_time(e1))
            keep = out  # noqa: F841  (previous result alive during the next call, like a caller's loop)
    cu = n * D**3 * a.steps
    for name, ts in times.items():
        print(json.dumps({'variant': name, 'window': os.environ.get('LNX_T64_WINDOW', 'default'), 'min_ms': min(ts), 'median_ms': pystat.median(ts),
                          'max_ms': max(ts), 'best_cell_updates_per_s': cu / (min(ts) * 1e-3), 'reps': len(ts)}), flush=True)


if __name__ == '__main__':
    main()
