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
    ap.add_argument('--no-early-stop', action='store_true', help='simulate every step of every world (the reference behaviour; default: early stop on)')
    ap.add_argument('--config', default='1c1k', choices=['1c1k', '3c6k'])
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
    if a.config == '3c6k':  # conf/config_qd_cmame_3c6k.yaml physics: 3 channels, 6 kernels, genotype (m, s, h) of every kernel
        import bench
        cfg['world_params']['nb_channels'] = 3
        cfg['kernels_params'] = bench.c3_kernels_params(1)[0]
        cfg['genotype'] = [{'key': f'kernels_params.{k}.{f}', 'domain': d, 'type': 'float'}
                           for k in range(6) for f, d in (('gf_params.0', [.1, .5]), ('gf_params.1', [.01, .1]), ('h', [.1, 1.]))]
        params = torch.rand(a.inds, 18, generator=g).tolist()
    eval_fn = qd.build_eval_lenia_config_mem_optimized_fn(cfg, device='cuda:0', early_stop=not a.no_early_stop)

    def generation():
        inds = [lenia.LeniaIndividual(copy.deepcopy(cfg), k, p) for k, p in zip(key.split(a.inds), params)]
        out = eval_fn(inds)
        torch.cuda.synchronize()
        return out

    generation()
    generation()
    walls = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = generation()
        walls.append(time.perf_counter() - t0)
    dt = sorted(walls)[len(walls) // 2]
    # where the host time of a generation goes (same calls as eval_fn, timed one by one with a device sync after each)
    from leniax_b200 import helpers, kernels, runner, statistics
    inds = [lenia.LeniaIndividual(copy.deepcopy(cfg), k, p) for k, p in zip(key.split(a.inds), params)]
    tt = [time.perf_counter()]
    rng_key, dyn = qd.get_dynamic_args(cfg, inds, True, device='cuda:0')
    torch.cuda.synchronize()
    tt.append(time.perf_counter())
    wp = cfg['world_params']
    ufn = helpers.build_update_fn(dyn[1][0].shape, kernels.get_kernels_and_mapping(copy.deepcopy(cfg['kernels_params']), [128, 128], wp['nb_channels'], wp['R'], device='cuda:0')[1])
    sfn = statistics.build_compute_stats_fn(wp, cfg['render_params'])
    tt[-1] = time.perf_counter()
    stats, _ = runner.run_scan_mem_optimized(rng_key, *dyn, a.steps, wp['R'], ufn, sfn, early_stop=not a.no_early_stop)
    tt.append(time.perf_counter())
    torch.cuda.synchronize()
    tt.append(time.perf_counter())
    qd.update_individuals(inds, stats, 1.)
    torch.cuda.synchronize()
    tt.append(time.perf_counter())
    phases = dict(zip(('get_dynamic_args (kernels + initial states, synced)', 'run_scan_mem_optimized host side (returns asynchronously)',
                       'wait for the scan', 'update_individuals (summary kernel + D2H + host loop)'), [1e3 * (b - a_) for a_, b in zip(tt, tt[1:])]))
    print('walls of 5 generations (ms):', [round(1e3 * w, 1) for w in walls], '| phases (ms):', {k: round(v, 2) for k, v in phases.items()})
    # device time of the scan kernel alone inside one generation (CUDA events around lnx_run_scan, via the profiler-free route: the
    # kernel is the only launch longer than a millisecond, so the difference wall - host shows up as the share below)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        generation()
    ev = [(e.key, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
    scan_ms = sum(ms for k, ms in ev if 'lnx_world128' in k or 'lnx::t' in k)
    all_ms = sum(ms for _, ms in ev)
    cu = a.inds * a.inits * 128 * 128 * a.steps
    print('one generation (%s, early stop %s): %d individuals x %d inits x %d steps: %.1f ms wall, scan kernel %.1f ms = %.1f %% of the wall time '
          '(all device work %.1f ms, %d other launches), %.3g cell-updates/s end to end; fitness %s' %
          (a.config, 'off' if a.no_early_stop else 'on', a.inds, a.inits, a.steps, dt * 1e3, scan_ms, 100 * scan_ms / (dt * 1e3), all_ms,
           sum(1 for k, _ in ev if 'lnx_world128' not in k), cu / dt, [o.fitness for o in out][:6]))
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        generation()
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
