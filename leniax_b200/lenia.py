"""LeniaIndividual: genotype -> configuration (reference: leniax/lenia.py:8-159)."""
import copy
from typing import Dict, List, Tuple

from . import utils as leniax_utils


class LeniaIndividual(object):
    """A Lenia individual used by QD algorithms (leniax/lenia.py:8-78)."""
    fitness: float
    features: List[float]

    def __init__(self, config: Dict, rng_key, params: List = []):
        self.qd_config = copy.deepcopy(config)
        self.rng_key = rng_key
        self.params = params
        self.fitness = 0.
        self.features = []
        if 'genotype' in self.qd_config:
            for gene in self.get_genotype():  # genotype keys must address existing values
                leniax_utils.get_param(self.qd_config, gene['key'])

    def set_init_props(self, rng_key, best_init_idxs: List[int]):
        self.qd_config['algo']['init_rng_key'] = rng_key.tolist() if hasattr(rng_key, 'tolist') else rng_key
        self.qd_config['algo']['best_init_idxs'] = best_init_idxs

    def set_cells(self, cells: str):
        self.qd_config['run_params']['cells'] = cells

    def set_init_cells(self, init_cells: str):
        self.qd_config['run_params']['init_cells'] = init_cells

    def get_config(self, read_only: bool = False) -> Dict:
        """lenia.py:56-78.  ``read_only=True`` (callers inside this package that only read the result, once per individual and
        generation): the sub-trees the genotype cannot touch are shared with ``self.qd_config`` instead of deep-copied."""
        if 'genotype' not in self.qd_config:
            return self.qd_config
        genotype = self.get_genotype()
        raw_values = [round(float(v), 8) for v in self.params]  # lenia.py:66
        assert len(raw_values) == len(genotype)
        return update_config(self.qd_config, get_update_config(genotype, raw_values), share_untouched=read_only)

    def get_genotype(self):
        return self.qd_config['genotype']


def update_config(config, to_update, share_untouched: bool = False):  # lenia.py:81-98
    if share_untouched:
        new_config = dict(config)
        for key in ('kernels_params', 'world_params'):
            if key in new_config:
                new_config[key] = copy.deepcopy(new_config[key])
    else:
        new_config = copy.deepcopy(config)
    if 'kernels_params' in to_update:
        for i, kernel in enumerate(to_update['kernels_params']):
            new_config['kernels_params'][i].update(kernel)
    if 'world_params' in to_update:
        new_config['world_params'].update(to_update['world_params'])
    return new_config


def linear_scale(raw_value: float, domain: Tuple[float, float]) -> float:  # lenia.py:131-143
    return domain[0] + (domain[1] - domain[0]) * raw_value


def log_scale(raw_value: float, domain: Tuple[float, float]) -> float:  # lenia.py:146-159
    return domain[0] * (domain[1] / domain[0])**raw_value


def get_update_config(genotype, raw_values: List) -> Dict:  # lenia.py:101-128
    to_update: Dict = {}
    for gene, raw in zip(genotype, raw_values):
        domain, kind = gene['domain'], gene['type']
        if kind == 'float':
            val = float(linear_scale(raw, domain))
        elif kind == 'int':
            val = int(linear_scale(raw, (domain[0], domain[1] + 1)) - 0.5)
        elif kind == 'choice':
            val = domain[int(linear_scale(raw, (0, len(domain))) - 0.5)]
        else:
            raise ValueError(f"type {kind} unknown")
        leniax_utils.set_param(to_update, gene['key'], val)
    return to_update
