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

    def __call__(self, state, K):
        """The potential alone, ``[N, K, *dims]`` (what the reference's ``get_potential_fn(state, K)`` returns): one
        ``update`` with identity growth and zero weights, of which only the potential is kept."""
        C = int(state.shape[1])
        nk = len(self.tc_indices) if self.tc_indices is not None else self.nb_slots
        return update(None, state, K, torch.zeros((nk, 2)), torch.zeros((C, nk)), 0., self, FieldFn(('identity', ) * nk, False),
                      get_state_simple)[2]


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
        max_k = pf.max_k_per_channel or pf.nb_slots // nb_channels  # the conv path learns C from the state
        c_in = tuple(s // max_k for s in slots)
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
    sfn = _resolve_state_fn(get_state_fn)
    dev = engine.require_cuda_device(state.device if isinstance(state, torch.Tensor) and state.is_cuda else None)
    state_t = engine.as_device_tensor(state, torch.float32, dev)
    N, C = state_t.shape[0], state_t.shape[1]
    world_size = tuple(state_t.shape[2:])
    ufn = UpdateFn(get_potential_fn, get_field_fn, sfn)
    slots, c_in, gf_ids = ufn.kernel_layout(C)
    if not get_potential_fn.fft:
        return update_conv(state_t, K, gf_params, kernels_weight_per_channel, dt, ufn)
    from . import kernels as _kernels
    if not _kernels.is_pow2_world(world_size):
        # the FFT engines need powers of two; the reference's fftn does not (core.py:81): same potential by direct convolution
        Kt = engine.as_device_tensor(K, torch.complex64, dev)
        taps = _kernels.spatial_from_spectrum(Kt, get_potential_fn.nb_slots, world_size)
        import dataclasses
        conv_ufn = UpdateFn(dataclasses.replace(get_potential_fn, fft=False), get_field_fn, sfn)
        return update_conv(state_t, taps, gf_params, kernels_weight_per_channel, dt, conv_ufn)
    dt_t = engine.as_device_tensor(dt, torch.float32, dev).reshape(-1)[:1]
    plan = engine.Plan.get(world_size=world_size, nb_channels=C, slots=slots, c_in=c_in, gf_ids=gf_ids,
                           nb_slots=get_potential_fn.nb_slots, state_fn=sfn.slug, weighted_average=get_field_fn.average,
                           R=1.0, stats_dt=1.0, device=dev)
    Kt = engine.as_device_tensor(K, torch.complex64, dev).reshape((1, get_potential_fn.nb_slots) + world_size)
    return plan.update(state_t, Kt, engine.as_device_tensor(gf_params, torch.float32, dev).reshape(len(slots), 2),
                       engine.as_device_tensor(kernels_weight_per_channel, torch.float32, dev).reshape(C, len(slots)), dt_t.contiguous())


def update_conv(state: torch.Tensor, K, gf_params, kernels_weight_per_channel, dt, ufn: UpdateFn
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``core.update`` with the direct-convolution potential (core.py:105-146, ``fft=False``): ``lnx_update_conv``.

    ``K`` is the reference's cropped kernel tensor ``[C * max_k, 1, kh, kw]`` (kernels.py:116-117, 153-156).  Any 2-D world
    size.  This is the cross-check path of the reference, not the throughput path.
    """
    import ctypes

    from . import _lib
    dev = state.device
    f32 = torch.float32
    N, C = state.shape[0], state.shape[1]
    world_size = tuple(state.shape[2:])
    if len(world_size) != 2:
        raise NotImplementedError('the direct-convolution potential is 2-D only, as in the reference (core.py:136)')
    slots, c_in, gf_ids = ufn.kernel_layout(C)
    Kt = engine.as_device_tensor(K, f32, dev)
    if Kt.dim() != 4 or Kt.shape[1] != 1 or Kt.shape[0] != ufn.get_potential_fn.nb_slots:
        raise ValueError(f'fft=False expects K of shape [C * max_k, 1, kh, kw], got {tuple(Kt.shape)}')
    kh, kw = int(Kt.shape[2]), int(Kt.shape[3])
    d = _lib.LnxDesc()
    d.nb_dims = 2
    d.dims[0], d.dims[1] = int(world_size[0]), int(world_size[1])
    d.nb_channels, d.nb_kernels, d.nb_slots = C, len(slots), ufn.get_potential_fn.nb_slots
    for k in range(len(slots)):
        d.slot[k], d.c_in[k], d.gf_id[k], d.c_out[k] = int(slots[k]), int(c_in[k]), int(gf_ids[k]), _lib.LNX_COUT_ANY
    d.state_fn = engine.STATE_FN_IDS[ufn.get_state_fn.slug]
    d.weighted_average = 1 if ufn.get_field_fn.average else 0
    d.R, d.stats_dt = 1.0, 1.0
    gf = engine.as_device_tensor(gf_params, f32, dev).reshape(len(slots), 2)
    w = engine.as_device_tensor(kernels_weight_per_channel, f32, dev).reshape(C, len(slots))
    dt_f = float(dt.reshape(-1)[0].item()) if isinstance(dt, torch.Tensor) else float(dt)
    new_state, field = torch.empty_like(state), torch.empty_like(state)
    potential = torch.empty((N, len(slots)) + world_size, dtype=f32, device=dev)
    lib = _lib.load_library()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.lnx_update_conv(ctypes.byref(d), N, kh, kw, state.data_ptr(), Kt.data_ptr(), gf.data_ptr(), w.data_ptr(), dt_f,
                                       new_state.data_ptr(), field.data_ptr(), potential.data_ptr(), stream))
    return new_state, field, potential
