"""Leniax core simulation functions — B200 descriptors (reference: leniax/core.py).

In the reference ``update`` is a jitted function parameterised by three traced callables
(``get_potential_fn``, ``get_field_fn``, ``get_state_fn``, core.py:13-49).  Here those three become small immutable
descriptors (what the CUDA kernels need to know: FFT/true-channel indices, growth-function enums + mean/sum, the
state-update variant) and ``update`` runs one fused step of ``lnx_run_scan``.  Passing an arbitrary Python callable
raises ``NotImplementedError`` — there is no CPU fallback.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import engine
from . import growth_functions as gfs


@dataclass(frozen=True)
class PotentialFn:
    """get_potential_fft bound to ``tc_indices`` (core.py:52-102, helpers.py:430-488)."""
    tc_indices: Optional[Tuple[int, ...]]
    nb_slots: int
    max_k_per_channel: int
    fft: bool = True
    channel_first: bool = True


@dataclass(frozen=True)
class FieldFn:
    """get_field bound to the growth functions and the weighted mean/sum (core.py:163-242, helpers.py:491-515)."""
    gf_slugs: Tuple[str, ...]
    average: bool = True


@dataclass(frozen=True)
class StateFn:
    """get_state / get_state_v2 / get_state_simple (core.py:245-319)."""
    slug: str

    def __call__(self, rng_key, state, field, dt):
        raise NotImplementedError('state updates run inside the fused CUDA step; use leniax_b200.core.update')


get_state = StateFn('v1')
get_state_v2 = StateFn('v2')
get_state_simple = StateFn('simple')

register = {'v1': get_state, 'v2': get_state_v2, 'simple': get_state_simple}  # core.py:322-326


def _resolve_state_fn(fn) -> StateFn:
    if isinstance(fn, StateFn):
        return fn
    if isinstance(fn, str) and fn in register:
        return register[fn]
    raise NotImplementedError(f'state function {fn!r} cannot be fused; supported: {sorted(register)}')


@dataclass(frozen=True)
class UpdateFn:
    """What ``helpers.build_update_fn`` returns: ``functools.partial(core.update, ...)`` in the reference."""
    get_potential_fn: PotentialFn
    get_field_fn: FieldFn
    get_state_fn: StateFn

    def __call__(self, rng_key, state, K, gf_params, kernels_weight_per_channel, dt):
        return update(rng_key, state, K, gf_params, kernels_weight_per_channel, dt, self.get_potential_fn, self.get_field_fn,
                      self.get_state_fn)

    # --- what the engine needs ---
    def kernel_layout(self, nb_channels: int):
        pf = self.get_potential_fn
        slots = tuple(pf.tc_indices) if pf.tc_indices is not None else tuple(range(pf.nb_slots))
        c_in = tuple(s // pf.max_k_per_channel for s in slots)
        gf_ids = tuple(gfs.resolve(s).gf_id for s in self.get_field_fn.gf_slugs)
        if len(gf_ids) != len(slots):
            raise ValueError(f'{len(gf_ids)} growth functions for {len(slots)} kernels')
        if max(c_in) >= nb_channels:
            raise ValueError('kernel input channel out of range')
        return slots, c_in, gf_ids


def update(rng_key, state, K, gf_params, kernels_weight_per_channel, dt, get_potential_fn, get_field_fn, get_state_fn
           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Update the cells state (core.py:13-49).  Returns ``(state, field, potential)``.

    ``state`` ``[N, C, H, W]``, ``K`` ``[1, C, max_k, H, W]`` complex64, ``gf_params`` ``[K, 2]``,
    ``kernels_weight_per_channel`` ``[C, K]``, ``dt`` scalar.  ``rng_key`` is unused, as in the reference.
    """
    if not isinstance(get_potential_fn, PotentialFn) or not isinstance(get_field_fn, FieldFn):
        raise NotImplementedError(
            'core.update needs the PotentialFn / FieldFn descriptors built by leniax_b200.helpers '
            '(arbitrary Python callables cannot be fused into the CUDA step; there is no CPU fallback)'
        )
    if not get_potential_fn.fft:
        raise NotImplementedError('the direct-convolution potential (fft=False, core.py:105-146) is not built; use fft=True')
    sfn = _resolve_state_fn(get_state_fn)
    dev = engine.require_cuda_device(state.device if isinstance(state, torch.Tensor) and state.is_cuda else None)
    state_t = engine.as_device_tensor(state, torch.float32, dev)
    N, C = state_t.shape[0], state_t.shape[1]
    world_size = tuple(state_t.shape[2:])
    ufn = UpdateFn(get_potential_fn, get_field_fn, sfn)
    slots, c_in, gf_ids = ufn.kernel_layout(C)
    dt_t = engine.as_device_tensor(dt, torch.float32, dev).reshape(-1)[:1]
    plan = engine.Plan.get(world_size=world_size, nb_channels=C, slots=slots, c_in=c_in, gf_ids=gf_ids,
                           nb_slots=get_potential_fn.nb_slots, state_fn=sfn.slug, weighted_average=get_field_fn.average,
                           R=1.0, stats_dt=1.0, device=dev)
    Kt = engine.as_device_tensor(K, torch.complex64, dev).reshape((1, get_potential_fn.nb_slots) + world_size)
    res = plan.run_scan(state_t[None], Kt, engine.as_device_tensor(gf_params, torch.float32, dev)[None],
                        engine.as_device_tensor(kernels_weight_per_channel, torch.float32, dev)[None], dt_t, 1,
                        keep_trajectory=True)
    return res['final_cells'][0], res['field'][0, 0], res['potential'][0, 0]
