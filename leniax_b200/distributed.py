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


def local_rectangles(n_sols: int, n_init: int, rank: int, world_size: int) -> List[Tuple[int, int, int, int]]:
    """This rank's slice of the flattened world axis as at most three ``(sol_start, sol_stop, init_start, init_stop)`` rectangles
    (a partial head solution, the whole solutions in between, a partial tail solution): each rectangle is ONE scan launch."""
    rects: List[Tuple[int, int, int, int]] = []
    for sol, i0, i1 in local_pieces(n_sols, n_init, rank, world_size):
        whole = i0 == 0 and i1 == n_init
        if whole and rects and rects[-1][2:] == (0, n_init) and rects[-1][1] == sol:
            rects[-1] = (rects[-1][0], sol + 1, 0, n_init)
        else:
            rects.append((sol, sol + 1, i0, i1))
    return rects


def _gather_device(fallback: torch.device, group) -> torch.device:
    """Device of the all-gather buffers: NCCL moves CUDA tensors only, so a rank without local worlds (or with host-side inputs)
    must still build its padded block on its own GPU."""
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return fallback


def _all_gather_rows(local: torch.Tensor, counts: List[int], group) -> torch.Tensor:
    """Concatenation over the ranks of ``local [counts[rank], n_cols]`` (ragged: padded to the longest, one all_gather)."""
    max_len = max(counts)
    padded = torch.zeros((max_len, local.shape[1]), dtype=torch.float32, device=local.device)
    padded[:local.shape[0]] = local
    gathered = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(gathered, padded, group=group)
    return torch.cat([g[:c] for g, c in zip(gathered, counts)])


def run_scan_mem_optimized_sharded(rng_key, cells0, K, gf_params, kernels_weight_per_channel, T, max_run_iter: int, R: float,
                                   update_fn, compute_stats_fn, group: Optional[dist.ProcessGroup] = None,
                                   local_run: Callable = None, early_stop: bool = False, sharded_inputs: Optional[str] = None
                                   ) -> Tuple[torch.Tensor, List[str], Dict[str, torch.Tensor]]:
    """Batched scan over all ranks of ``group``; returns ``(summary [N_sols, N_init, 1 + n_keys] on every rank, key order, this
    rank's full statistics of its last launch)``.

    ``sharded_inputs=None``: every rank passes the same full arguments (``cells0 [N_sols, N_init, C, *dims]`` ..., on the host or on
    the device); a rank touches only its own slice of them — its worlds run in at most three launches (``local_rectangles``), a
    slice that covers whole solutions (BASELINE configs[2]: 2 solutions per GPU) or part of one (configs[1]: 512 of 4096 inits
    per GPU) in ONE.  ``sharded_inputs='sols'``: the arguments hold only this rank's solutions (ranks in order; counts may
    differ), ``'inits'``: only this rank's slice of the initialisation axis of every solution — no rank ever holds the whole batch.
    ``local_run`` defaults to ``runner.run_scan_mem_optimized`` (tests inject a CPU stand-in)."""
    if local_run is None:
        local_run = lambda *a: leniax_runner.run_scan_mem_optimized(*a, early_stop=early_stop, return_final_cells=False)  # noqa: E731
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    n_sols, n_init = cells0.shape[0], cells0.shape[1]
    n_cols = 1 + len(leniax_qd.STAT_KEYS_FOR_SUMMARY)
    keys: Optional[List[str]] = None
    last_stats: Dict[str, torch.Tensor] = {}

    def run_rect(s0, s1, i0, i1):
        nonlocal keys, last_stats
        stats, _ = local_run(rng_key, cells0[s0:s1, i0:i1], K[s0:s1], gf_params[s0:s1], kernels_weight_per_channel[s0:s1], T[s0:s1],
                             max_run_iter, R, update_fn, compute_stats_fn)
        block, keys = leniax_qd.summarize_stats(stats)
        last_stats = stats
        return block  # [s1 - s0, i1 - i0, 1 + n_keys]

    if sharded_inputs in ('sols', 'inits'):
        block = run_rect(0, n_sols, 0, n_init) if n_sols * n_init > 0 else None
        axis = 0 if sharded_inputs == 'sols' else 1
        if world == 1:
            return block, list(keys), last_stats
        dev = _gather_device(block.device if block is not None else cells0.device, group)
        mine = torch.tensor([n_sols, n_init], dtype=torch.int64, device=dev)
        shapes = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(shapes, mine, group=group)
        shapes = [tuple(int(v) for v in t.tolist()) for t in shapes]
        other = {shp[1 - axis] for shp in shapes if shp[0] * shp[1] > 0}
        if len(other) != 1:
            raise ValueError(f'sharded_inputs={sharded_inputs!r}: the ranks disagree on the unsharded axis: {shapes}')
        other = other.pop()
        n_cols = block.shape[-1] if block is not None else n_cols
        # rows of the gather = this rank's block with the sharded axis leading
        local = (block if axis == 0 else block.transpose(0, 1)).reshape(-1, other * n_cols).to(dev) if block is not None \
            else torch.zeros((0, other * n_cols), device=dev)
        rows = _all_gather_rows(local.contiguous(), [shp[axis] for shp in shapes], group).reshape(-1, other, n_cols)
        summary = rows if axis == 0 else rows.transpose(0, 1).contiguous()
        return summary, list(keys) if keys is not None else list(leniax_qd.STAT_KEYS_FOR_SUMMARY), last_stats
    if sharded_inputs is not None:
        raise ValueError(f"sharded_inputs must be None, 'sols' or 'inits', got {sharded_inputs!r}")

    blocks = [run_rect(*r).reshape(-1, n_cols) for r in local_rectangles(n_sols, n_init, rank, world)]
    device = _gather_device(blocks[0].device if blocks else cells0.device, group)
    local = torch.cat(blocks).to(device) if blocks else torch.zeros((0, n_cols), device=device)
    if world > 1:
        sizes = [shard_range(n_sols * n_init, r, world) for r in range(world)]
        local = _all_gather_rows(local, [b - a for a, b in sizes], group)
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
            tmp = dict(ind.get_config(read_only=True))
            tmp['behaviours'] = {k: float(s[i, best, 1 + j]) for j, k in enumerate(keys)}
            ind.features = [leniax_utils.get_param(tmp, key) for key in ind.qd_config['phenotype']]
    return inds
