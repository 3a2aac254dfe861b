"""Host-side helpers the hot path needs (reference: leniax/utils.py:27-263) — config fix-ups and cell merging.

Hydra/OmegaConf are not required: ``get_container`` accepts a plain dict (e.g. ``yaml.safe_load`` of a leniax YAML file).
"""
import copy
from fractions import Fraction
from typing import Any, Dict, List

import torch


def st2fracs2float(st: str) -> List[float]:  # utils.py:214-225
    return [float(Fraction(s)) for s in st.split(',')]


def update_config_v1_v2(config: Dict) -> Dict:  # utils.py:91-150
    config['version'] = 2
    old_gf = {0: 'poly_quad4', 1: 'gaussian', 2: 'gaussian_target', 3: 'step'}
    old_kf = {0: 'poly_quad', 1: 'gauss_bump', 2: 'step', 3: 'staircase', 4: 'gauss'}
    new = []
    for kp in config['kernels_params']['k']:
        bs = st2fracs2float(kp['b']) if isinstance(kp['b'], str) else kp['b']
        new.append({
            'k_slug': 'circle_2d',
            'k_params': [kp['r'] if 'r' in kp else 1., bs],
            'kf_slug': old_kf[kp['k_id']],
            'kf_params': [kp['q']],
            'gf_slug': old_gf[kp['gf_id']],
            'gf_params': [kp['m'], kp['s']],
            'h': kp['h'],
            'c_in': kp['c_in'],
            'c_out': kp['c_out'],
        })
    config['kernels_params'] = new
    for gen in config.get('genotype', []) or []:
        if 'kernels_params.k' in gen['key']:
            gen['key'] = gen['key'].replace('kernels_params.k', 'kernels_params').replace('.m', '.gf_params.0').replace('.s', '.gf_params.1')
    for i, ph in enumerate(config.get('phenotype', []) or []):
        if 'kernels_params.k' in ph:
            config['phenotype'][i] = ph.replace('kernels_params.k', 'kernels_params').replace('.m', '.gf_params.0').replace('.s', '.gf_params.1')
    slug = config['world_params'].get('get_state_fn_slug', 'v1')
    config.setdefault('algo', {})
    config['algo']['init_slug'] = 'perlin' if slug == 'v1' else ('perlin_local' if slug == 'v2' else config['algo'].get('init_slug', 'perlin'))
    config['algo']['init_param'] = []
    return config


def get_container(config: Dict, main_path: str = '') -> Dict:
    """utils.py:27-88 for a plain dict: fills world_size / pixel_size / scale defaults and upgrades v1 configs."""
    config = copy.deepcopy(dict(config))
    config.pop('hydra', None)
    wp, rp = config['world_params'], config['render_params']
    if rp.get('pixel_size', 'MISSING') == 'MISSING':
        rp['pixel_size'] = 2**rp.get('pixel_size_power2', 0)
    if rp.get('world_size', 'MISSING') == 'MISSING':
        rp['world_size'] = [2**rp['size_power2']] * wp['nb_dims']
    config['main_path'] = main_path
    wp.setdefault('scale', 1.)
    config.setdefault('algo', {})
    config.setdefault('other', {}).setdefault('log_level', 20)
    if 'update_fn_version' in wp:
        wp['get_state_fn_slug'] = wp.pop('update_fn_version')
    if config.get('version', 1) == 1:
        config = update_config_v1_v2(config)
    return config


def load_config(path: str) -> Dict:
    import os

    import yaml
    with open(path, 'r', encoding='utf-8') as f:
        return get_container(yaml.safe_load(f), os.path.dirname(os.path.abspath(path)))


def get_param(dic: Dict, key_string: str) -> Any:  # utils.py:153-177
    for key in key_string.split('.'):
        dic = dic[int(key)] if key.isdigit() else dic[key]
    if isinstance(dic, torch.Tensor):
        return float(dic.float().mean())
    if isinstance(dic, (int, float)):
        return float(dic)
    if isinstance(dic, str):
        return dic
    raise ValueError(f"dic type {type(dic)} not supported")


def set_param(dic: Dict, key_string: str, value: Any):  # utils.py:180-211
    keys = key_string.split('.')
    for i, key in enumerate(keys[:-1]):
        if key.isdigit():
            if not isinstance(dic, list):
                raise Exception('This key should be an array')
            idx = int(key)
            while len(dic) < idx + 1:
                dic.append({})
            dic = dic[idx]
        else:
            dic = dic.setdefault(key, [] if keys[i + 1].isdigit() else {})
    if keys[-1].isdigit():
        idx = int(keys[-1])
        while len(dic) < idx + 1:
            dic.append(None)
        dic[idx] = value
    else:
        dic[keys[-1]] = value


def merge_cells(cells: torch.Tensor, other_cells: torch.Tensor, offset: List[int] = None) -> torch.Tensor:  # utils.py:231-263
    assert cells.dim() == other_cells.dim()
    assert cells.shape[0] == other_cells.shape[0]
    offset = offset or [0] * cells.dim()
    assert len(offset) == cells.dim()
    pads: List[int] = []
    for i in reversed(range(cells.dim())):
        start = int(max((cells.shape[i] - other_cells.shape[i]) // 2 + offset[i], 0))
        pads += [start, int(cells.shape[i] - other_cells.shape[i] - start)]
    return cells + torch.nn.functional.pad(other_cells.to(cells.dtype), pads)


def zoom_nearest(x: torch.Tensor, scale: float) -> torch.Tensor:
    """``scipy.ndimage.zoom(x, scale, order=0)`` (the call of leniax/helpers.py:61): output size ``round(n * scale)`` per axis,
    output index ``i`` reads input index ``floor(i * (n_in - 1) / (n_out - 1) + 0.5)``.  Checked bit for bit against scipy in
    tests/test_host_logic.py."""
    out_shape = [int(round(n * scale)) for n in x.shape]
    idx = []
    for n_in, n_out in zip(x.shape, out_shape):
        c = torch.arange(n_out, dtype=torch.float64) * ((n_in - 1) / (n_out - 1)) if n_out > 1 else torch.zeros(1, dtype=torch.float64)
        idx.append(torch.floor(c + 0.5).long().clamp(0, n_in - 1).to(x.device))
    return x[torch.meshgrid(*idx, indexing='ij')]
