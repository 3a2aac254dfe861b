#!/usr/bin/env python
"""Small runs of the resident kernels for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Fused TMEM kernel (1 channel, 1 kernel), generic TMEM kernel (orbium-scutium: 2 channels, 2 kernels), early stop on."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leniax_b200 import helpers, runner, statistics, utils  # noqa: E402

DEV = 'cuda:0'
for name, n, steps in (('orbium-test', 5, 40), ('orbium-scutium-test', 3, 36)):
    cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', name + '.yaml'))
    cells, K, mapping = helpers.init(copy.deepcopy(cfg), device=DEV)
    wp = cfg['world_params']
    ufn = helpers.build_update_fn(K.shape, mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), True)
    sfn = statistics.build_compute_stats_fn(wp, cfg['render_params'])
    gf, w = mapping.get_gf_params(DEV), mapping.get_kernels_weight_per_channel(DEV)
    cells0 = torch.stack([torch.roll(cells[0], (7 * i, 3 * i), dims=(1, 2)) for i in range(n)])[None]
    T = torch.tensor([float(wp['T'])], device=DEV)
    for early in (False, True):
        stats, final = runner.run_scan_mem_optimized(None, cells0, K[None], gf[None], w[None], T, steps, wp['R'], ufn, sfn, early_stop=early)
        torch.cuda.synchronize()
        print(name, 'early' if early else 'full', 'N =', stats['N'][0].tolist(), 'mass[-1] =', [round(float(x), 5) for x in stats['mass'][0, -1]])
print('done')
