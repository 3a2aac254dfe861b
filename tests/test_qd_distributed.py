"""QD consumer + multi-rank sharding logic (CPU: gloo, world_size 2; the simulation itself is a CPU stand-in built on the
oracle because the CUDA path needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leniax_b200 import distributed as lnx_dist
from leniax_b200 import initializations, lenia, qd, utils
from oracle import lenia_oracle as lo


def _fake_stats(n_sols, T, n_init, seed=0):
    rng = np.random.default_rng(seed)
    stats = {k: torch.from_numpy(rng.random((n_sols, T, n_init)).astype(np.float32)) for k in qd.STAT_KEYS_FOR_SUMMARY}
    stats['channel_mass'] = torch.from_numpy(rng.random((n_sols, T, n_init, 1)).astype(np.float32))
    stats['N'] = torch.from_numpy(rng.integers(0, T + 1, size=(n_sols, n_init)).astype(np.float32))
    return stats


def test_get_config_kat():  # reference tests/test_qd.py:14-50
    config = {
        'kernels_params': [{'r': 1, 'b': "1", 'm': 0.17, 's': 0.015, 'h': 1, 'k_id': 0, 'gf_id': 0, 'c_in': 0, 'c_out': 0}],
        'genotype': [{'key': 'kernels_params.0.m', 'domain': [0., .5], 'type': 'float'},
                     {'key': 'kernels_params.0.s', 'domain': [0., 2.], 'type': 'float'}]
    }
    ind = lenia.LeniaIndividual(config, initializations.RngKey(1), [0.7, 0.3])
    new = ind.get_config()
    assert new['kernels_params'][0]['m'] == pytest.approx(0.35) and new['kernels_params'][0]['s'] == pytest.approx(0.6)
    assert new['genotype'] == config['genotype']


def test_update_individuals_fitness_kat():  # reference tests/test_qd.py:90-109
    cfg = {'kernels_params': [], 'algo': {}, 'run_params': {}}
    inds = [lenia.LeniaIndividual(cfg, initializations.RngKey(1), []), lenia.LeniaIndividual(cfg, initializations.RngKey(1), [])]
    new = qd.update_individuals(inds, {'N': torch.tensor([[1., 2., 3.], [1., 3., 4.]])})
    assert new[0].fitness == 3 and new[1].fitness == 4
    assert new[0].qd_config['algo']['best_init_idxs'] == [2]


def test_summary_matches_reference_consumer():
    """summarize_stats == qd.py:168-186 (argmax init, mean of rows [ns-128, ns), ns = max(N, 128))."""
    stats = _fake_stats(3, 300, 5)
    block, keys = qd.summarize_stats(stats)
    fitness, best, behaviours = lo.behaviours_of({k: v.numpy() for k, v in stats.items() if k != 'channel_mass'})
    for i in range(3):
        assert float(block[i, :, 0].max()) == fitness[i]
        for j, k in enumerate(keys):
            assert float(block[i, best[i], 1 + j]) == pytest.approx(float(behaviours[i][k]), rel=1e-6)
    cfg = {'kernels_params': [], 'algo': {}, 'run_params': {}, 'phenotype': ['behaviours.mass_density', 'behaviours.mass_speed']}
    inds = [lenia.LeniaIndividual(cfg, initializations.RngKey(i), []) for i in range(3)]
    inds = qd.update_individuals(inds, stats)
    for i in range(3):
        assert inds[i].fitness == fitness[i]
        assert inds[i].features == pytest.approx([float(behaviours[i]['mass_density']), float(behaviours[i]['mass_speed'])], rel=1e-6)


def test_perlin_init_properties():
    key = initializations.RngKey(3)
    k2, cells = initializations.perlin(key, 16, [128, 128], 13, [0.15, 0.015], device='cpu')
    assert cells.shape == (16, 1, 128, 128) and k2.seed != key.seed
    scal = [0.15 + i / 16 * (0.45 - 0.15) for i in range(16)]  # initializations.py:57-60
    np.testing.assert_allclose(cells.amax(dim=(1, 2, 3)).numpy(), scal, atol=1e-4)
    assert float(cells.min()) == 0.
    q = cells * 12543
    assert float((q - q.round()).abs().max()) < 1e-2  # make_array_compressible (loader.py:16-30)
    again = initializations.perlin(key, 16, [128, 128], 13, [0.15, 0.015], device='cpu')[1]
    assert torch.equal(cells, again)  # counter-based key: reproducible
    _, u = initializations.random_uniform(key, 4, [128, 128], 13, [0.15, 0.015], device='cpu')
    assert u.shape == (4, 128, 128) and float(u[0].max()) <= 0.4 + 1e-6


def test_shard_ranges_cover_everything():
    for n, w in [(4096, 8), (2048, 8), (10, 4), (3, 8), (1, 2)]:
        got = [lnx_dist.shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n and all(a[1] == b[0] for a, b in zip(got, got[1:]))
        assert max(b - a for a, b in got) - min(b - a for a, b in got) <= 1
    pieces = [p for r in range(4) for p in lnx_dist.local_pieces(3, 10, r, 4)]
    covered = sorted((s, i) for s, a, b in pieces for i in range(a, b))
    assert covered == [(s, i) for s in range(3) for i in range(10)]
    # rectangles = launches: BASELINE configs[2] (16 solutions x 128 inits over 8 ranks) and configs[1] (1 x 4096) are one launch per rank
    assert lnx_dist.local_rectangles(16, 128, 3, 8) == [(6, 8, 0, 128)]
    assert lnx_dist.local_rectangles(1, 4096, 7, 8) == [(0, 1, 3584, 4096)]
    for n_sols, n_init, world in [(3, 10, 4), (5, 7, 3), (16, 128, 8), (2, 3, 8)]:
        rects = [r for rk in range(world) for r in lnx_dist.local_rectangles(n_sols, n_init, rk, world)]
        assert all(len(lnx_dist.local_rectangles(n_sols, n_init, rk, world)) <= 3 for rk in range(world))
        cov = sorted((s, i) for s0, s1, a, b in rects for s in range(s0, s1) for i in range(a, b))
        assert cov == [(s, i) for s in range(n_sols) for i in range(n_init)]


# ---- 2-rank gloo run: each rank simulates its slice with a CPU stand-in (oracle), one all_gather, same result ----
def _oracle_local_run(rng_key, cells0, K, gf_params, W, T, max_run_iter, R, update_fn, compute_stats_fn):
    upd, sfn = update_fn, compute_stats_fn
    stats, final = lo.run_scan_mem_optimized(cells0.numpy(), K.numpy(), gf_params.numpy(), W.numpy(), T.numpy(), max_run_iter, upd, sfn)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in stats.items()}, torch.from_numpy(final)


def _worker(rank, world, port, golden_dir, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cfg = lo.load_yaml_config(os.path.join(golden_dir, 'orbium-test.yaml'))
        cells, K, mapping = lo.init(cfg)
        rng = np.random.default_rng(0)
        n_sols, n_init, steps = 2, 3, 20
        worlds = np.stack([np.stack([np.roll(cells[0], (int(rng.integers(128)), int(rng.integers(128))), axis=(1, 2)) * (0.2 if (s + i) % 3 == 0 else 1.)
                                     for i in range(n_init)]) for s in range(n_sols)]).astype(np.float32)
        args = (torch.from_numpy(worlds), torch.from_numpy(np.stack([K] * n_sols)), torch.from_numpy(np.stack([mapping.get_gf_params()] * n_sols)),
                torch.from_numpy(np.stack([mapping.get_kernels_weight_per_channel()] * n_sols)), torch.tensor([10., 10.]))
        upd = lo.build_update_fn(mapping)
        sfn = lo.build_compute_stats_fn(cfg['world_params'], cfg['render_params'])
        summary, keys, _ = lnx_dist.run_scan_mem_optimized_sharded(None, *args, steps, 13, upd, sfn, local_run=_oracle_local_run)
        # the same batch with every rank holding only its own part: one solution each / a ragged split of the initialisations
        s0, s1 = lnx_dist.shard_range(n_sols, rank, world)
        by_sols, _, _ = lnx_dist.run_scan_mem_optimized_sharded(None, *[a[s0:s1] for a in args], steps, 13, upd, sfn,
                                                                local_run=_oracle_local_run, sharded_inputs='sols')
        i0, i1 = lnx_dist.shard_range(n_init, rank, world)  # 3 inits over 2 ranks: 2 + 1
        by_inits, _, _ = lnx_dist.run_scan_mem_optimized_sharded(None, args[0][:, i0:i1], *args[1:], steps, 13, upd, sfn,
                                                                 local_run=_oracle_local_run, sharded_inputs='inits')
        torch.save({'summary': summary, 'keys': keys, 'pieces': lnx_dist.local_pieces(n_sols, n_init, rank, world), 'by_sols': by_sols,
                    'by_inits': by_inits}, os.path.join(out_dir, f'rank{rank}.pt'))
        if rank == 0:
            full, _ = _oracle_local_run(None, *args, steps, 13, upd, sfn)
            torch.save(qd.summarize_stats(full)[0], os.path.join(out_dir, 'single.pt'))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process(golden_dir, tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, golden_dir, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / 'rank0.pt'), torch.load(tmp_path / 'rank1.pt')
    single = torch.load(tmp_path / 'single.pt')
    assert torch.equal(r0['summary'], r1['summary'])  # every rank holds the gathered block
    assert torch.equal(r0['summary'], single)  # and it equals the unsharded computation bit for bit
    assert r0['pieces'] != r1['pieces'] and len(r0['keys']) == 11
    for k in ('by_sols', 'by_inits'):  # pre-sharded inputs: same gathered block on both ranks, equal to the unsharded one
        assert torch.equal(r0[k], r1[k]) and torch.equal(r0[k], single), k
