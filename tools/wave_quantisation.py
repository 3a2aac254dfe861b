#!/usr/bin/env python
"""Time of the headline scan (Orbium 1c1k 128x128, 1024 steps, all statistics) against the number of worlds on ONE GPU: the staircase the
strong-scaling split of BASELINE configs[1] walks down (4096 worlds / N GPUs on 2 x 148 resident CTA slots).  One JSON line per size."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
for n in (sms // 2, sms, 2 * sms, 2 * sms + 1, 3 * sms, 512, 4 * sms, 4 * sms + 1, 1024, 2048, 4096):
    wl = bench.Workload('B', 'strong', 0, 1, dev, n_override=n)
    for _ in range(2):
        wl.step(wl.dev_cells)
    ms = wl.kernel_only_ms(3)
    waves = n / (2 * sms)
    print(json.dumps({'worlds': n, 'kernel_ms': round(ms, 3), 'slots': 2 * sms, 'waves': round(waves, 3),
                      'cell_updates_per_s': n * 128 * 128 * wl.sim_steps / (ms * 1e-3),
                      'efficiency_vs_4096': None}), flush=True)
