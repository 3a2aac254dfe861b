"""Helper functions gluing the hot path together (reference: leniax/helpers.py:35-128, 130-190, 401-515)."""
import copy
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import core as leniax_core
from . import kernels as leniax_kernels
from . import loader as leniax_loader
from . import runner as leniax_runner
from . import statistics as leniax_stat
from . import utils as leniax_utils
from .growth_functions import resolve as resolve_gf


def create_init_cells(world_size: List[int], nb_channels: int, other_cells: Union[torch.Tensor, List] = [],
                      offsets: List[List[int]] = []) -> torch.Tensor:
    """helpers.py:91-128."""
    if isinstance(other_cells, list):
        cells = torch.zeros([nb_channels] + list(world_size))
        use_off = len(offsets) == len(other_cells)
        for i, c in enumerate(other_cells):
            if len(c) != 0:
                cells = leniax_utils.merge_cells(cells, torch.as_tensor(c), offsets[i] if use_off else None)
        return cells[None]
    if isinstance(other_cells, (torch.Tensor, np.ndarray)):
        return torch.as_tensor(other_cells).to(torch.float32)
    raise ValueError(f'Don\'t know how to handle {type(other_cells)}')


def init(config: Dict, use_init_cells: bool = True, fft: bool = True, device=None
         ) -> Tuple[torch.Tensor, torch.Tensor, leniax_kernels.KernelMapping]:
    """Initial state, kernels and mapping (helpers.py:35-88)."""
    wp = config['world_params']
    nb_dims, nb_channels = wp['nb_dims'], wp['nb_channels']
    world_size = list(config['render_params']['world_size'])
    assert len(world_size) == nb_dims
    assert nb_channels > 0
    raw_cells = leniax_loader.load_raw_cells(config, use_init_cells)
    scale = wp.get('scale', 1.)
    if scale != 1.:  # helpers.py:58-66: nearest-neighbour zoom of every channel; the new R is used by the kernels and the statistics
        raw_cells = torch.stack([leniax_utils.zoom_nearest(raw_cells[i], scale) for i in range(nb_channels)]).to(torch.float32)
        wp['R'] *= scale
    if raw_cells.dim() > 1 + nb_dims:
        init_cells = create_init_cells(world_size, nb_channels, raw_cells)
    else:
        init_cells = create_init_cells(world_size, nb_channels, [raw_cells])
    K, mapping = leniax_kernels.get_kernels_and_mapping(config['kernels_params'], world_size, nb_channels, wp['R'], fft, device=device)
    return init_cells.to(K.device), K, mapping


def build_get_potential_fn(kernel_shape: Tuple[int, ...], true_channels: Optional[List[bool]] = None, fft: bool = True,
                           channel_first: bool = True) -> leniax_core.PotentialFn:
    """helpers.py:430-488.  ``kernel_shape`` is ``K.shape`` = ``[1, C, max_k, H, W]`` for the FFT path and
    ``[C * max_k, 1, kh, kw]`` for the direct-convolution path (``fft=False``)."""
    if not channel_first:
        raise NotImplementedError('channel_first=False layouts are not built')
    tc = tuple(i for i, t in enumerate(true_channels) if t) if true_channels is not None else None
    if not fft:
        # direct convolution (helpers.py:464-488): K is [C * max_k, 1, kh, kw]; C is only known once the state is seen
        if len(kernel_shape) != 4:
            raise NotImplementedError('the direct-convolution potential is 2-D only, as in the reference (core.py:136)')
        return leniax_core.PotentialFn(tc, int(kernel_shape[0]), 0, False, True)
    C, max_k = int(kernel_shape[1]), int(kernel_shape[2])
    return leniax_core.PotentialFn(tc, C * max_k, max_k, True, True)


def build_get_field_fn(cin_gfs: List[List[str]], average: bool = True) -> leniax_core.FieldFn:
    """helpers.py:491-515."""
    slugs = tuple(resolve_gf(s).slug for per_channel in cin_gfs for s in per_channel)
    return leniax_core.FieldFn(slugs, bool(average))


def build_update_fn(kernel_shape: Tuple[int, ...], mapping: leniax_kernels.KernelMapping, get_state_fn_slug: str = 'v1',
                    average_weight: bool = True, fft: bool = True) -> leniax_core.UpdateFn:
    """helpers.py:401-427."""
    return leniax_core.UpdateFn(
        build_get_potential_fn(kernel_shape, mapping.true_channels, fft),
        build_get_field_fn(mapping.cin_gfs, average_weight),
        leniax_core._resolve_state_fn(get_state_fn_slug),
    )


def init_and_run(rng_key, config: Dict, use_init_cells: bool = True, with_jit: bool = True, fft: bool = True,
                 stat_trunc: bool = False, device=None):
    """Initialise and simulate a Lenia configuration (helpers.py:130-190)."""
    config = copy.deepcopy(config)
    cells, K, mapping = init(config, use_init_cells, fft, device=device)
    dev = K.device
    gf_params = mapping.get_gf_params(dev)
    weights = mapping.get_kernels_weight_per_channel(dev)
    wp = config['world_params']
    R = wp['R']
    T = torch.tensor(wp['T'], dtype=torch.float32, device=dev)
    max_run_iter = config['run_params']['max_run_iter']
    update_fn = build_update_fn(K.shape, mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), fft)
    stats_fn = leniax_stat.build_compute_stats_fn(wp, config['render_params'])
    if with_jit:
        all_cells, all_fields, all_potentials, stats = leniax_runner.run_scan(rng_key, cells, K, gf_params, weights, T, max_run_iter,
                                                                             R, update_fn, stats_fn)
    else:
        all_cells, all_fields, all_potentials, stats = leniax_runner.run(rng_key, cells, K, gf_params, weights, T, max_run_iter, R,
                                                                        update_fn, stats_fn, stat_trunc)
    stats = {k: v.squeeze() for k, v in stats.items()}
    if stat_trunc:
        n = int(stats['N'])
        all_cells, all_fields, all_potentials = all_cells[:n], all_fields[:n], all_potentials[:n]
    return all_cells, all_fields, all_potentials, stats


def multi_init_and_run(rng_key, main_config: Dict, configs: List[Dict], use_init_cells: bool = True, fft: bool = True, device=None):
    """Several configurations of the same structure in one launch (helpers.py:192-237: ``vmap`` of ``run_scan`` over
    cells / K / gf_params / weights / T).  Returns ``(cells, field, potential, stats)`` with a leading configuration axis."""
    main_config = copy.deepcopy(main_config)
    wp = main_config['world_params']
    _, K0, mapping0 = init(main_config, use_init_cells, fft, device=device)
    dev = K0.device
    update_fn = build_update_fn(K0.shape, mapping0, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), fft)
    stats_fn = leniax_stat.build_compute_stats_fn(wp, main_config['render_params'])
    cells_l, K_l, gf_l, w_l, T_l = [], [], [], [], []
    for config in configs:
        cells, K, mapping = init(copy.deepcopy(config), use_init_cells, fft, device=dev)
        cells_l.append(cells)
        K_l.append(K)
        gf_l.append(mapping.get_gf_params(dev))
        w_l.append(mapping.get_kernels_weight_per_channel(dev))
        T_l.append(float(config['world_params']['T']))
    res = leniax_runner._scan(torch.stack(cells_l), torch.stack(K_l), torch.stack(gf_l), torch.stack(w_l),
                              torch.tensor(T_l, dtype=torch.float32, device=dev), main_config['run_params']['max_run_iter'], update_fn,
                              stats_fn, batched=True, keep_trajectory=True)
    stats = {k: v.squeeze() for k, v in res['stats'].items()}
    return res['cells'], res['field'], res['potential'], stats


def search_for_init(rng_key, config: Dict, fft: bool = True, device=None) -> Tuple[Dict, int]:
    """Search for a stable initial state (helpers.py:318-397).

    The reference simulates the ``nb_init_search`` initialisations one after the other with ``run_scan`` and stops at the
    first one that survives ``max_run_iter`` steps; the best run is the first one reaching the running maximum of ``N``.
    Here all initialisations run at once (statistics only), the loop's stopping index and best index are read off ``N``,
    and only the best initialisation is re-simulated with its trajectory.  Returns ``(best_run, i)`` like the reference.
    """
    from . import initializations as leniax_init
    wp = config['world_params']
    nb_channels, R = wp['nb_channels'], wp['R']
    world_size = list(config['render_params']['world_size'])
    kernels_params = config['kernels_params']
    nb_init_search = config['run_params']['nb_init_search']
    max_run_iter = config['run_params']['max_run_iter']
    K, mapping = leniax_kernels.get_kernels_and_mapping(kernels_params, world_size, nb_channels, R, fft, device=device)
    dev = K.device
    gf_params, weights = mapping.get_gf_params(dev), mapping.get_kernels_weight_per_channel(dev)
    T = torch.tensor(float(wp['T']), dtype=torch.float32, device=dev)
    update_fn = build_update_fn(K.shape, mapping, wp.get('get_state_fn_slug', 'v1'), wp.get('weighted_average', True), fft)
    stats_fn = leniax_stat.build_compute_stats_fn(wp, config['render_params'])
    rng_key, noises = leniax_init.register[config['algo']['init_slug']](rng_key, nb_channels * nb_init_search, world_size, R,
                                                                         kernels_params[0]['gf_params'], device=dev)
    all_cells0 = noises.reshape([nb_init_search, 1, nb_channels] + world_size).to(torch.float32)  # one world per "run_scan"
    stats, _ = leniax_runner.run_scan_mem_optimized(rng_key, all_cells0.reshape([1, nb_init_search, nb_channels] + world_size), K[None],
                                                    gf_params[None], weights[None], T.reshape(1), max_run_iter, R, update_fn, stats_fn)
    N = stats['N'][0].cpu()
    survivors = torch.nonzero(N >= max_run_iter).flatten()
    i_stop = int(survivors[0]) if len(survivors) else nb_init_search - 1  # where the reference's loop breaks
    best = int(torch.argmax(N[:i_stop + 1]))  # first index reaching the running maximum
    best_run: Dict = {}
    if float(N[best]) > 0:  # the reference keeps {} when no run survives its first step (current_max < N is strict)
        all_cells, _, _, all_stats = leniax_runner.run_scan(rng_key, all_cells0[best], K, gf_params, weights, T, max_run_iter, R,
                                                             update_fn, stats_fn)
        best_run = {'N': all_stats['N'], 'all_cells': all_cells, 'all_stats': all_stats}
    return best_run, i_stop


def search_for_mutation(rng_key, config: Dict, nb_scale_for_stability: int = 1, use_init_cells: bool = True, fft: bool = True,
                        mutation_rate: float = 1e-5, device=None) -> Tuple[Dict, int]:
    """Search for a stable mutation (helpers.py:240-315): every gene gets Gaussian noise of std ``mutation_rate``, the mutated
    configuration is simulated at ``nb_scale_for_stability`` scales (world size and ``scale`` doubled each time) and the one
    with the most steps done over all scales wins; stops at the first one surviving everywhere.  ``rng_key`` is a
    ``leniax_b200.initializations.RngKey`` (the Gaussian draws are torch's, not jax.random's: parity of the stream is
    unpinned)."""
    world_size = config['render_params']['world_size']
    nb_mut_search = config['run_params']['nb_mut_search']
    max_run_iter = config['run_params']['max_run_iter']
    best_run: Dict = {}
    current_max = 0
    nb_genes = len(config['genotype'])
    subkeys = rng_key.split(nb_mut_search * nb_genes + 1)[1:]
    i = 0
    for i in range(nb_mut_search):
        copied_config = copy.deepcopy(config)
        for gene_i, gene in enumerate(config['genotype']):
            val = leniax_utils.get_param(copied_config, gene['key'])
            noise = torch.randn((), generator=subkeys[i * nb_genes + gene_i].generator('cpu'), dtype=torch.float32)
            leniax_utils.set_param(copied_config, gene['key'], float(val + float(noise) * mutation_rate))
        total_iter_done, nb_iter_done = 0, 0
        for scale_power in range(nb_scale_for_stability):
            scaled_config = copy.deepcopy(copied_config)
            scaled_config['render_params']['world_size'] = [ws * 2**scale_power for ws in world_size]
            scaled_config['world_params']['scale'] = 2**scale_power
            all_cells, _, _, stats_dict = init_and_run(rng_key, scaled_config, use_init_cells=use_init_cells, with_jit=True, fft=fft,
                                                       device=device)
            all_cells = all_cells[:, 0]
            n = int(stats_dict['N'])
            nb_iter_done = max(nb_iter_done, n)
            total_iter_done += n
        if current_max < total_iter_done:
            current_max = total_iter_done
            best_run = {'N': nb_iter_done, 'all_cells': all_cells, 'all_stats': stats_dict, 'config': copied_config}
        if total_iter_done >= max_run_iter * nb_scale_for_stability:
            break
    return best_run, i
