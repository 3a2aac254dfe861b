"""Multi-GPU sharding of the batched scan (replaces ``runner.run_scan_mem_optimized_pmap``,
leniax/runner.py:218-268, which maps replicas with ``jax.pmap`` and gathers on the host).

One process per GPU (``torch.distributed``, NCCL on GPUs, gloo in the CPU tests).  Worlds are independent, so rank ``g``
simulates a contiguous slice of the flattened ``(N_sols, N_init)`` world axis with no data-path collective; a single
``all_gather`` of the ``[worlds, 1 + 11]`` fitness/behaviour block (what ``qd.update_individuals`` consumes) closes a call.
"""
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import qd as leniax_qd
from . import runner as leniax_runner


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice of ``range(n_items)`` owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_pieces(n_sols: int, n_init: int, rank: int, world_size: int) -> List[Tuple[int, int, int]]:
    """``(sol, init_start, init_stop)`` pieces covering this rank's slice of the flattened world axis."""
    w0, w1 = shard_range(n_sols * n_init, rank, world_size)
    pieces = []
    w = w0
    while w < w1:
        sol, i0 = divmod(w, n_init)
        i1 = min(n_init, i0 + (w1 - w))
        pieces.append((sol, i0, i1))
        w += i1 - i0
    return pieces


def run_scan_mem_optimized_sharded(rng_key, cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float,
                                   update_fn, compute_stats_fn, group: Optional[dist.ProcessGroup] = None,
                                   local_run: Callable = None, early_stop: bool = False
                                   ) -> Tuple[torch.Tensor, List[str], Dict[str, torch.Tensor]]:
    """Every rank passes the same full arguments (``cells0 [N_sols, N_init, C, H, W]`` …); each simulates its slice.

    Returns ``(summary [N_sols, N_init, 1 + n_keys] on every rank, key order, this rank's full statistics of its last
    piece)``.  ``local_run`` defaults to ``runner.run_scan_mem_optimized`` (tests inject a CPU stand-in)."""
    if local_run is None:
        local_run = lambda *a: leniax_runner.run_scan_mem_optimized(*a, early_stop=early_stop)  # noqa: E731
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    n_sols, n_init = cells0.shape[0], cells0.shape[1]
    blocks, keys, last_stats = [], None, {}
    for sol, i0, i1 in local_pieces(n_sols, n_init, rank, world):
        stats, _ = local_run(rng_key, cells0[sol:sol + 1, i0:i1], K[sol:sol + 1], gf_params[sol:sol + 1],
                             kernels_weight_per_channel[sol:sol + 1], T[sol:sol + 1], max_run_iter, R, update_fn, compute_stats_fn)
        block, keys = leniax_qd.summarize_stats(stats)
        blocks.append(block[0])  # [i1 - i0, 1 + n_keys]
        last_stats = stats
    n_cols = 1 + len(leniax_qd.STAT_KEYS_FOR_SUMMARY) if keys is None else 1 + len(keys)
    device = blocks[0].device if blocks else cells0.device
    local = torch.cat(blocks) if blocks else torch.zeros((0, n_cols), device=device)
    if world > 1:
        sizes = [shard_range(n_sols * n_init, r, world) for r in range(world)]
        max_len = max(b - a for a, b in sizes)
        padded = torch.zeros((max_len, n_cols), dtype=torch.float32, device=device)
        padded[:local.shape[0]] = local
        gathered = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded, group=group)
        local = torch.cat([g[:b - a] for g, (a, b) in zip(gathered, sizes)])
    if keys is None:
        keys = list(leniax_qd.STAT_KEYS_FOR_SUMMARY)
    return local.reshape(n_sols, n_init, n_cols), keys, last_stats


def update_individuals_from_summary(inds, summary: torch.Tensor, keys: List[str], fitness_coef=1.):
    """``qd.update_individuals`` (leniax/qd.py:150-188) fed by the all-gathered summary block."""
    from . import utils as leniax_utils
    s = summary.cpu()
    Ns = s[..., 0]
    for i, ind in enumerate(inds):
        mx = Ns[i].max()
        best = int(torch.argmax(Ns[i]))
        ind.set_init_props(ind.rng_key, torch.nonzero(Ns[i] == mx).flatten().tolist())
        ind.fitness = float(fitness_coef * mx)
        if 'phenotype' in ind.qd_config:
            tmp = ind.get_config()
            tmp['behaviours'] = {k: float(s[i, best, 1 + j]) for j, k in enumerate(keys)}
            ind.features = [leniax_utils.get_param(tmp, key) for key in ind.qd_config['phenotype']]
    return inds
