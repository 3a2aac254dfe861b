#!/usr/bin/env python
"""A/B timing of BASELINE config D (one 2048^2 world, R = 52, 256 steps) that survives a noisy box: the variants of the step loop alternate
inside ONE process and the minimum / median over many repetitions are reported (like tools/ab_config_e.py).

    python tools/ab_config_d.py [--reps 20]

Variants: steps per captured CUDA graph (LNX_GRAPH_UNROLL) x programmatic dependent launches between the kernels of a graph (LNX_T2K_PDL);
the library reads both switches at every scan.
"""
import argparse
import json
import os
import statistics as pystat
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from leniax_b200 import helpers, kernels, runner, statistics  # noqa: E402

DEV = 'cuda:0'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--steps', type=int, default=256)
    ap.add_argument('--unrolls', default='1,4,16')
    ap.add_argument('--burst', type=int, default=1, help='calls per timed region, enqueued without a synchronisation in between')
    a = ap.parse_args()
    size, R = 2048, 52
    K, mapping = kernels.get_kernels_and_mapping(bench.ORBIUM_KP, [size, size], 1, R, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': R, 'T': 10}, {'world_size': [size, size]})
    cells = torch.from_numpy(bench.d_world_numpy())[None, None, None].to(DEV)
    gf, w = mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None]
    T = torch.tensor([10.], device=DEV)

    def run():
        return runner.run_scan_mem_optimized(None, cells, K[None], gf, w, T, a.steps, R, ufn, sfn)

    variants = [(int(u), p) for u in a.unrolls.split(',') for p in (0, 1)]
    times = {v: [] for v in variants}
    mass = {}
    keep = None
    for rep in range(a.reps + 2):
        for v in variants:
            os.environ['LNX_GRAPH_UNROLL'], os.environ['LNX_T2K_PDL'] = str(v[0]), str(v[1])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.burst):
                out = run()
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                times[v].append(e0.elapsed_time(e1) / a.burst)
            mass[v] = float(out[0]['mass'][0, -1, 0])
            keep = out  # noqa: F841  (previous result alive during the next call, like a caller's loop)
    cu = size * size * a.steps
    for v, ts in times.items():
        print(json.dumps({'steps_per_graph': v[0], 'pdl': v[1], 'min_ms': min(ts), 'median_ms': pystat.median(ts), 'max_ms': max(ts),
                          'best_cell_updates_per_s': cu / (min(ts) * 1e-3), 'last_mass': mass[v], 'reps': len(ts), 'burst': a.burst}), flush=True)


if __name__ == '__main__':
    main()
