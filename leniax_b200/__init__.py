"""leniax_b200 — B200-native engine for the Lenia simulation hot path.

Drop-in replacement for the hot path of morgangiraud/leniax (``core.update``, ``runner.run_scan`` /
``run_scan_mem_optimized``, ``statistics.build_compute_stats_fn``, ``kernels.get_kernels_and_mapping`` and the QD batch
evaluation path).  Host code is Python; all arithmetic of the path runs in hand-written sm_100a CUDA kernels reached
through the C ABI declared in ``include/leniax_b200.h`` (``libleniax_b200.so``).  PyTorch tensors are used only as device
buffers.  There is no CPU fallback: every entry point raises if the library or a B200-class GPU is missing.
"""
from . import (constant, core, distributed, growth_functions, helpers, initializations, kernel_functions, kernels, lenia, loader, qd,  # noqa: F401
               runner, statistics, utils)
from ._lib import LeniaxB200Error, library_path, load_library  # noqa: F401

__version__ = '0.1.0'
