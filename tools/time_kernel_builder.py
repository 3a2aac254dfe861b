#!/usr/bin/env python
"""Cost of kernels.get_kernels_and_mapping per QD individual (3 channels, 6 kernels) with and without the kernel-shape cache."""
import time, copy, torch, sys
sys.path.insert(0, ".")
from leniax_b200 import kernels
pairs = [(0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 0)]
base = [dict(k_slug="circle_2d", k_params=[1., [1.]], kf_slug="poly_quad", kf_params=[4], gf_slug="poly_quad4", gf_params=[.17, .015], h=1., c_in=p[0], c_out=p[1]) for p in pairs]
def run(n):
    torch.cuda.synchronize(); t=time.time()
    for i in range(n):
        kp = copy.deepcopy(base); kp[0]["gf_params"][0] = .1 + i * 1e-3
        K, m = kernels.get_kernels_and_mapping(kp, [128, 128], 3, 13, device="cuda:0")
    torch.cuda.synchronize(); return (time.time() - t) / n * 1e3
kernels._K_CACHE_MAX = 0; kernels._K_CACHE.clear(); run(2); a = run(16)
kernels._K_CACHE_MAX = 64; run(2); b = run(16)
print("get_kernels_and_mapping 3c6k per individual: %.2f ms uncached, %.2f ms cached" % (a, b))
