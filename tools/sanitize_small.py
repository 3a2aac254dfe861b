#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py [resident|tiled|setup]
resident: lnx_world128_tm (1 channel, 1 kernel), lnx_world128_gen2 (orbium-scutium: 2 channels, 2 kernels; 3c6k), lnx_world128_gen_tm
          (same worlds, LNX_RUN_GENERIC_1CTA, and a trajectory scan), early stop on and off
tiled:    2048^2 four-step engine with both rows kernels (a 2048^2 world is big: 3 steps); 64^3: whole-scan kernel, half-line kernels per
          pass and step, round-1 thread-per-line kernels, generic tiled passes; generic tiled passes (256^2)
anysize:  a 100 x 120 world: taps from the spectrum, lnx_update_conv, any-size lnx_compute_stats
setup:    lnx_rasterize_kernels, lnx_kernel_spectrum (2-D, 3-D), lnx_init_perlin(_seeded), lnx_init_uniform, lnx_summarize_stats"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from leniax_b200 import helpers, initializations, kernels, qd, runner, statistics, utils  # noqa: E402

DEV = 'cuda:0'
what = sys.argv[1] if len(sys.argv) > 1 else 'resident'


def orbium(size, R):
    kp = copy.deepcopy(bench.ORBIUM_KP)
    K, mapping = kernels.get_kernels_and_mapping(kp, list(size), 1, R, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': R, 'T': 10}, {'world_size': list(size)})
    return K, mapping, ufn, sfn


if what == 'resident':
    for name, n, steps in (('orbium-test', 5, 40), ('orbium-scutium-test', 3, 36)):
        cfg = utils.load_config(os.path.join(ROOT, 'tests', 'golden', name + '.yaml'))
        cells, K, mapping = helpers.init(copy.deepcopy(cfg), device=DEV)
        wp = cfg['world_params']
        ufn = helpers.build_update_fn(K.shape, mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), True)
        sfn = statistics.build_compute_stats_fn(wp, cfg['render_params'])
        gf, w = mapping.get_gf_params(DEV), mapping.get_kernels_weight_per_channel(DEV)
        cells0 = torch.stack([torch.roll(cells[0], (7 * i, 3 * i), dims=(1, 2)) for i in range(n)])[None]
        T = torch.tensor([float(wp['T'])], device=DEV)
        for one_cta in (False, True):
            runner.GENERIC_1CTA = one_cta
            for early in (False, True):
                stats, final = runner.run_scan_mem_optimized(None, cells0, K[None], gf[None], w[None], T, steps, wp['R'], ufn, sfn, early_stop=early)
                torch.cuda.synchronize()
                print(name, 'gen_tm' if one_cta else 'default', 'early' if early else 'full', 'N =', stats['N'][0].tolist())
        runner.GENERIC_1CTA = False
        runner.run_scan(None, cells0[0, :2], K, gf, w, T[0], 6, wp['R'], ufn, sfn)  # trajectory scan (gen_tm)
        torch.cuda.synchronize()
    kps = bench.c3_kernels_params(2)
    Ks, maps = kernels.get_kernels_and_mapping_batch(copy.deepcopy(kps), [128, 128], 3, 13, device=DEV)
    _, cells = initializations.perlin_batch([initializations.RngKey(i) for i in range(2)], 9, [128, 128], 13, [kp[0]['gf_params'] for kp in kps], device=DEV)
    ufn = helpers.build_update_fn(Ks[0].shape, maps[0])
    sfn = statistics.build_compute_stats_fn({'R': 13, 'T': 10}, {'world_size': [128, 128]})
    gf, w = torch.stack([m.get_gf_params(DEV) for m in maps]), torch.stack([m.get_kernels_weight_per_channel(DEV) for m in maps])
    stats, _ = runner.run_scan_mem_optimized(None, cells.reshape(2, 3, 3, 128, 128), Ks, gf, w, torch.full((2, ), 10., device=DEV), 34, 13, ufn, sfn)
    torch.cuda.synchronize()
    print('3c6k gen2 N =', stats['N'].tolist())
elif what == 'tiled':
    K, mapping, ufn, sfn = orbium([2048, 2048], 52)
    cells = torch.from_numpy(bench.d_world_numpy())[None, None, None].to(DEV)
    for real_rows in (False, True):
        runner.T2K_REAL_ROWS = real_rows
        stats, _ = runner.run_scan_mem_optimized(None, cells, K[None], mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None],
                                                 torch.tensor([10.], device=DEV), 3, 52, ufn, sfn)
        torch.cuda.synchronize()
        print('2048^2', 'real rows' if real_rows else 'row pairs', 'mass', stats['mass'][0, :, 0].tolist())
    runner.T2K_REAL_ROWS = False
    kern = torch.from_numpy(bench.sphere_kernel_numpy(13)).to(DEV)
    kp = [dict(bench.ORBIUM_KP[0], k_slug='raw', k_params=kern)]
    K, mapping = kernels.get_kernels_and_mapping(kp, [64, 64, 64], 1, 13, device=DEV)
    ufn = helpers.build_update_fn(K.shape, mapping)
    sfn = statistics.build_compute_stats_fn({'R': 13, 'T': 10}, {'world_size': [64, 64, 64]})
    _, cells = initializations.random_uniform(initializations.RngKey(5), 3, [64, 64, 64], 13, [.15, .015], device=DEV)
    for eng in ('whole_scan', 'stepwise', 'line64', 'generic'):
        runner.TILED_GENERIC, runner.T64_STEPWISE, runner.T64_LINE = eng == 'generic', eng == 'stepwise', eng == 'line64'
        stats, _ = runner.run_scan_mem_optimized(None, cells[None, :, None], K[None], mapping.get_gf_params(DEV)[None],
                                                 mapping.get_kernels_weight_per_channel(DEV)[None], torch.tensor([10.], device=DEV), 9, 13, ufn, sfn)
        torch.cuda.synchronize()
        print('64^3', eng, 'mass', [round(float(x), 4) for x in stats['mass'][0, -1]])
    runner.TILED_GENERIC = runner.T64_STEPWISE = runner.T64_LINE = False
    K, mapping, ufn, sfn = orbium([256, 256], 13)
    cells = (torch.rand((1, 2, 1, 256, 256), device=DEV) * .4)
    stats, _ = runner.run_scan_mem_optimized(None, cells, K[None], mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None],
                                             torch.tensor([10.], device=DEV), 9, 13, ufn, sfn)
    torch.cuda.synchronize()
    print('256^2 generic passes mass', [round(float(x), 4) for x in stats['mass'][0, -1]])
elif what == 'anysize':
    K, mapping, ufn, sfn = orbium([100, 120], 13)
    cells = torch.zeros((1, 2, 1, 100, 120), device=DEV)
    cells[0, :, 0, 20:60, 30:70] = torch.rand((2, 40, 40), device=DEV) * .6
    stats, _ = runner.run_scan_mem_optimized(None, cells, K[None], mapping.get_gf_params(DEV)[None], mapping.get_kernels_weight_per_channel(DEV)[None],
                                             torch.tensor([10.], device=DEV), 5, 13, ufn, sfn)
    torch.cuda.synchronize()
    print('100 x 120 mass', [round(float(x), 4) for x in stats['mass'][0, -1]])
else:
    kps = bench.c3_kernels_params(3)
    kps[1][0].update(k_slug='ellipse_2d', k_params=[1., [1., .5], .9, .6, .25])
    Ks, _ = kernels.get_kernels_and_mapping_batch(kps, [128, 128], 3, 13, device=DEV)
    K3 = kernels.kernel_spectrum(torch.from_numpy(bench.sphere_kernel_numpy(13)).to(DEV), [64, 64, 64])
    keys, cells = initializations.perlin_batch([initializations.RngKey(i) for i in range(3)], 5, [128, 128], 13, [[.15, .015]] * 3, device=DEV)
    _, one = initializations.perlin(initializations.RngKey(9), 4, [128, 128], 13, [.15, .015], device=DEV)
    _, uni = initializations.random_uniform(initializations.RngKey(9), 3, [32, 32, 32], 13, [.15, .015], device=DEV)
    stats = {k: torch.rand((2, 140, 7), device=DEV) for k in qd.STAT_KEYS_FOR_SUMMARY}
    stats['N'] = torch.randint(0, 141, (2, 7), device=DEV).float()
    block, _ = qd.summarize_stats(stats)
    torch.cuda.synchronize()
    print('setup kernels:', float(Ks.abs().sum()), float(K3.abs().sum()), float(cells.sum()), float(one.sum()), float(uni.sum()), float(block.sum()))
print('done')
