"""Growth-function registry (reference: leniax/growth_functions.py:256-264).

In the reference the registry maps a slug to a traced JAX callable.  Here it maps the same slugs to descriptors
carrying the enum the CUDA kernels switch on (``lnx_growth_fn`` in include/leniax_b200.h); the arithmetic itself lives
in ``csrc/lnx_step.cuh``.  A user-defined Python callable cannot be fused into the persistent kernel, so registering one
(as examples/cgol.py:31-56 does upstream) raises ``NotImplementedError`` when the update function is built.
"""
from dataclasses import dataclass
from typing import Dict


@dataclass(frozen=True)
class GrowthFunction:
    slug: str
    gf_id: int  # lnx_growth_fn
    nb_params: int = 2

    def __call__(self, params, X):
        raise NotImplementedError(
            f"growth function '{self.slug}' is evaluated inside the fused CUDA step (leniax_b200.core.update); "
            'it has no standalone host implementation'
        )


poly_quad4 = GrowthFunction('poly_quad4', 0)
gaussian = GrowthFunction('gaussian', 1)
gaussian_target = GrowthFunction('gaussian_target', 2)
step = GrowthFunction('step', 3)
staircase = GrowthFunction('staircase', 4)
triangle = GrowthFunction('triangle', 5)
identity = GrowthFunction('identity', 6)

register: Dict[str, GrowthFunction] = {
    'poly_quad4': poly_quad4,
    'gaussian': gaussian,
    'gaussian_target': gaussian_target,
    'step': step,
    'staircase': staircase,
    'triangle': triangle,
    'identity': identity,
}


def resolve(slug_or_fn) -> GrowthFunction:
    if isinstance(slug_or_fn, GrowthFunction):
        return slug_or_fn
    if isinstance(slug_or_fn, str) and isinstance(register.get(slug_or_fn), GrowthFunction):
        return register[slug_or_fn]
    raise NotImplementedError(
        f'growth function {slug_or_fn!r} is not one of the fused CUDA growth functions {sorted(register)}; '
        'arbitrary Python callables cannot run inside the persistent kernel and there is no CPU fallback'
    )
