"""QD batch-evaluation path (reference: leniax/qd.py:33-188).

Only the evaluation half is rebuilt: ``build_eval_lenia_config_mem_optimized_fn``, ``get_dynamic_args`` and
``update_individuals``.  The pyribs search loop, archives, plots and pickling of ``qd.py:191-633`` are unchanged Python
above this boundary and stay in the reference.
"""
from typing import Callable, Dict, List, Tuple

import torch

from . import helpers as leniax_helpers
from . import initializations as leniax_init
from . import kernels as leniax_kernels
from . import runner as leniax_runner
from . import utils as leniax_utils
from . import _lib
from .constant import NB_STATS_STEPS
from .lenia import LeniaIndividual
from .statistics import build_compute_stats_fn

STAT_KEYS_FOR_SUMMARY = ('mass', 'mass_volume', 'mass_density', 'growth', 'growth_volume', 'growth_density', 'mass_speed',
                         'mass_angle_speed', 'mass_growth_dist', 'inertia', 'potential_volume')


def build_eval_lenia_config_mem_optimized_fn(qd_config: Dict, fitness_coef: float = 1., fft: bool = True, device=None,
                                             early_stop: bool = False) -> Callable:
    """qd.py:33-77.  ``early_stop=True`` (extension) skips the steps a stopped world no longer needs."""
    max_run_iter = qd_config['run_params']['max_run_iter']
    world_params, render_params = qd_config['world_params'], qd_config['render_params']
    R = world_params['R']
    K, mapping = leniax_kernels.get_kernels_and_mapping(qd_config['kernels_params'], render_params['world_size'],
                                                        world_params['nb_channels'], R, fft, device=device)
    update_fn = leniax_helpers.build_update_fn(K.shape, mapping, world_params.get('get_state_fn_slug', 'v1'),
                                               world_params.get('weighted_average', True), fft)
    compute_stats_fn = build_compute_stats_fn(world_params, render_params)

    def eval_lenia_config_mem_optimized(leniax_sols: List[LeniaIndividual]) -> List[LeniaIndividual]:
        cfg = leniax_sols[0].qd_config
        rng_key, dynamic_args = get_dynamic_args(cfg, leniax_sols, fft, device=device)
        stats, _ = leniax_runner.run_scan_mem_optimized(rng_key, *dynamic_args, max_run_iter, R, update_fn, compute_stats_fn,
                                                        early_stop=early_stop)
        return update_individuals(leniax_sols, stats, fitness_coef)

    return eval_lenia_config_mem_optimized


def get_dynamic_args(qd_config: Dict, leniax_sols: List[LeniaIndividual], fft: bool = True, device=None
                     ) -> Tuple[object, Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]:
    """qd.py:80-147: per-individual kernels + initial states, stacked on a leading ``N_sols`` axis."""
    world_params = qd_config['world_params']
    nb_channels, R = world_params['nb_channels'], world_params['R']
    world_size = qd_config['render_params']['world_size']
    nb_init_search = qd_config['run_params']['nb_init_search']
    init_slug = qd_config['algo']['init_slug']
    cells0, Ks, gfs, ws, Ts = [], [], [], [], []
    rng_key = None
    for ind in leniax_sols:
        config = ind.get_config()
        kernels_params = config['kernels_params']
        K, mapping = leniax_kernels.get_kernels_and_mapping(kernels_params, world_size, nb_channels, R, fft, device=device)
        nb_init = nb_channels * nb_init_search
        rng_key, noises = leniax_init.register[init_slug](ind.rng_key, nb_init, world_size, R, kernels_params[0]['gf_params'],
                                                          device=K.device)
        cells0.append(noises.reshape([nb_init_search, nb_channels] + list(world_size)))
        Ks.append(K)
        gfs.append(mapping.get_gf_params(K.device))
        ws.append(mapping.get_kernels_weight_per_channel(K.device))
        Ts.append(float(config['world_params']['T']))
    dev = Ks[0].device
    return rng_key, (torch.stack(cells0), torch.stack(Ks), torch.stack(gfs), torch.stack(ws), torch.tensor(Ts, dtype=torch.float32, device=dev))


def summarize_stats(stats: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, List[str]]:
    """Per world: N and the mean of rows ``[ns-128, ns)`` (``ns = max(N, 128)``) of every scalar statistic — exactly what
    ``update_individuals`` reads (qd.py:181-185).  Returns ``[N_sols, N_init, 1 + n_keys]`` and the key order; this is the
    block all-gathered across GPUs."""
    keys = [k for k in stats if k not in ('N', 'channel_mass')]
    N = stats['N']  # [S, I]
    if N.is_cuda and tuple(keys) == tuple(_lib.STAT_KEYS) and all(stats[k].is_cuda and stats[k].dtype == torch.float32 for k in keys):
        import ctypes
        S, T, I = stats[keys[0]].shape
        planes = [stats[k].contiguous() for k in keys]
        ptrs = (ctypes.c_void_p * len(planes))(*[p.data_ptr() for p in planes])
        n_alive = N.contiguous().float()
        out = torch.empty((S, I, 1 + len(keys)), dtype=torch.float32, device=N.device)
        with torch.cuda.device(N.device):
            _lib.check(_lib.load_library().lnx_summarize_stats(ptrs, n_alive.data_ptr(), S, T, I, NB_STATS_STEPS, out.data_ptr(),
                                                               torch.cuda.current_stream().cuda_stream))
        return out, keys
    # host-side tensors (post-processing of stored statistics): same reduction with torch
    T = stats[keys[0]].shape[1]
    ns = torch.clamp(N.long(), min=min(NB_STATS_STEPS, T), max=T)
    lo = torch.clamp(ns - NB_STATS_STEPS, min=0)
    cols = [N]
    for k in keys:
        cs = torch.cat([torch.zeros_like(stats[k][:, :1], dtype=torch.float64), stats[k].double().cumsum(dim=1)], dim=1)  # [S, T+1, I]
        hi_v = torch.gather(cs, 1, ns[:, None, :]).squeeze(1)
        lo_v = torch.gather(cs, 1, lo[:, None, :]).squeeze(1)
        cols.append(((hi_v - lo_v) / (ns - lo).double()).float())
    return torch.stack(cols, dim=-1), keys


def update_individuals(inds: List[LeniaIndividual], stats: Dict[str, torch.Tensor], fitness_coef=1.) -> List[LeniaIndividual]:
    """qd.py:150-188: fitness = coef * max over inits of N; behaviours = mean of the last 128 rows before ``ns``."""
    Ns = stats['N']
    all_best = torch.argmax(Ns, dim=1)
    all_max = Ns.max(dim=1).values
    need_behaviours = any('phenotype' in ind.qd_config for ind in inds)
    block, keys = summarize_stats(stats) if need_behaviours and len(stats) > 1 else (None, [])
    Ns_h, best_h, max_h = Ns.cpu(), all_best.cpu(), all_max.cpu()
    block_h = block.cpu() if block is not None else None
    for i, ind in enumerate(inds):
        best_idxs = torch.nonzero(Ns_h[i] == max_h[i]).flatten().tolist()
        ind.set_init_props(ind.rng_key, best_idxs)
        ind.fitness = float(fitness_coef * max_h[i])
        if 'phenotype' in ind.qd_config:
            tmp = ind.get_config()
            tmp['behaviours'] = {k: float(block_h[i, int(best_h[i]), 1 + j]) for j, k in enumerate(keys)} if block_h is not None else {}
            ind.features = [leniax_utils.get_param(tmp, key) for key in ind.qd_config['phenotype']]
    return inds


def grid_archive_index(features: torch.Tensor, grid_shape: List[int], features_domain: List[List[float]]) -> torch.Tensor:
    """Cell of a pyribs ``GridArchive(grid_shape, features_domain)`` (examples/qd_cmame.py:60-64) a behaviour descriptor
    falls into: ``[..., D]`` float -> ``[..., D]`` int64.  Restates ``GridArchive.get_index`` of ribs 0.4.0 (the reference's
    pin, setup.py:21) — clip ``bc + eps`` to ``[lower, upper - eps]`` with eps = 1e-6, then
    ``int((bc - lower) / (upper - lower) * dims)``.  pyribs is not in this image: **parity unpinned** (SURVEY §8c)."""
    f = features.to(torch.float64)
    lower = torch.tensor([d[0] for d in features_domain], dtype=torch.float64, device=f.device)
    upper = torch.tensor([d[1] for d in features_domain], dtype=torch.float64, device=f.device)
    dims = torch.tensor(list(grid_shape), dtype=torch.float64, device=f.device)
    eps = 1e-6
    f = torch.minimum(torch.maximum(f + eps, lower), upper - eps)
    return ((f - lower) / (upper - lower) * dims).to(torch.int64)
