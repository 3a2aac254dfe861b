#!/usr/bin/env python
"""Per-call timing of BASELINE config D (one 2048x2048 world, 256 steps) through runner.run_scan_mem_optimized: CUDA-event time and
host wall time of each of 12 consecutive calls, then a cProfile of one call.  The call is short (about 10 ms of kernels), so host work
that is not hidden under the kernels shows up directly in the event time."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench_configs as bc  # noqa: E402
from leniax_b200 import runner  # noqa: E402

size, R, steps = 2048, 52, 256
K, mapping, ufn, sfn = bc.orbium_parts(size, R)
big = np.kron(bc.orbium_cells(), np.ones((4, 4), np.float32))
world = np.zeros((size, size), np.float32)
rng = np.random.default_rng(4)
for _ in range(16):
    y, x = rng.integers(0, size - big.shape[0], 2)
    world[y:y + big.shape[0], x:x + big.shape[1]] = np.maximum(world[y:y + big.shape[0], x:x + big.shape[1]], big)
cells = torch.from_numpy(world).to('cuda:0')[None, None, None]
args = (cells, K[None], mapping.get_gf_params('cuda:0')[None], mapping.get_kernels_weight_per_channel('cuda:0')[None], torch.full((1, ), 10., device='cuda:0'))
call = lambda: runner.run_scan_mem_optimized(None, *args, steps, R, ufn, sfn)  # noqa: E731
call()
torch.cuda.synchronize()
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    call()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'call {i:2d}: events {e0.elapsed_time(e1):7.2f} ms   host until return {1e3 * (t1 - t0):7.2f} ms   until synchronised {1e3 * (t2 - t0):7.2f} ms', flush=True)
pr = cProfile.Profile()
pr.enable()
call()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)


def back_to_back(label, fn, n=8):
    """n calls without a host synchronisation between them (what tools/bench_configs.py's timed() does)."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    host = []
    ev[0].record()
    for i in range(n):
        t0 = time.perf_counter()
        fn()
        host.append(1e3 * (time.perf_counter() - t0))
        ev[i + 1].record()
    torch.cuda.synchronize()
    print(label, 'events:', ' '.join(f'{ev[i].elapsed_time(ev[i + 1]):.1f}' for i in range(n)), '| host:', ' '.join(f'{h:.1f}' for h in host), flush=True)


back_to_back('back-to-back, same dt tensor      ', call)
back_to_back('back-to-back, same dt tensor      ', call)
args2 = args[:4]
back_to_back('back-to-back, dt tensor per call  ', lambda: runner.run_scan_mem_optimized(None, *args2, torch.tensor([10.], device='cuda:0'), steps, R, ufn, sfn))
keep = []
back_to_back('back-to-back, results kept alive  ', lambda: keep.append(call()))
