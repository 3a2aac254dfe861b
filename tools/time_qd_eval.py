#!/usr/bin/env python
"""Wall time of one QD generation through the reference-facing entry point (qd.build_eval_lenia_config_mem_optimized_fn,
leniax/qd.py:33-77): N individuals x nb_init_search perlin initialisations x max_run_iter steps, with a cProfile of the host side.

    python tools/time_qd_eval.py [--inds 16] [--inits 128] [--steps 1024] [--profile]
"""
import argparse
import cProfile
import copy
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leniax_b200 import initializations, lenia, qd, utils  # noqa: E402

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--inds', type=int, default=16)
    ap.add_argument('--inits', type=int, default=128)
    ap.add_argument('--steps', type=int, default=1024)
    ap.add_argument('--profile', action='store_true')
    a = ap.parse_args()
    cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', 'orbium-test.yaml'))
    cfg['run_params']['max_run_iter'] = a.steps
    cfg['run_params']['nb_init_search'] = a.inits
    cfg['algo']['init_slug'] = 'perlin'
    cfg['genotype'] = [{'key': 'kernels_params.0.gf_params.0', 'domain': [0.1, 0.3], 'type': 'float'},
                       {'key': 'kernels_params.0.gf_params.1', 'domain': [0.01, 0.04], 'type': 'float'}]
    cfg['phenotype'] = ['behaviours.mass_density', 'behaviours.mass_speed']
    key = initializations.RngKey(7)
    g = torch.Generator().manual_seed(0)
    params = torch.rand(a.inds, 2, generator=g).tolist()
    eval_fn = qd.build_eval_lenia_config_mem_optimized_fn(cfg, device='cuda:0')

    def generation():
        inds = [lenia.LeniaIndividual(copy.deepcopy(cfg), k, p) for k, p in zip(key.split(a.inds), params)]
        out = eval_fn(inds)
        torch.cuda.synchronize()
        return out

    generation()
    t0 = time.perf_counter()
    out = generation()
    dt = time.perf_counter() - t0
    cu = a.inds * a.inits * 128 * 128 * a.steps
    print('one generation: %d individuals x %d inits x %d steps: %.1f ms wall, %.3g cell-updates/s end to end; fitness %s' %
          (a.inds, a.inits, a.steps, dt * 1e3, cu / dt, [o.fitness for o in out][:6]))
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        generation()
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
