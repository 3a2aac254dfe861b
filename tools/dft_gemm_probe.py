#!/usr/bin/env python
"""DFT-as-GEMM on the tensor cores for the 128x128 potential (the variant BASELINE.json north_star says to use "only if measured to meet
tolerance").  potential = real(F^-1 (F x F^T . K) F^-T) with the 128-point DFT matrix F, as four complex matrix products per world on
torch.matmul (cuBLAS): fp32 reference, TF32 tensor cores, 3xTF32 (operands split in hi + lo TF32 parts, three products) and bf16x3.
Reports the L-inf error of one potential against fp64 (the fp32 FFT kernels sit at <= 3e-7; the state tolerance is 1e-5 after 64 steps and
the growth function multiplies potential errors by up to 100, so the potential must be good to a few 1e-7) and the GPU time per world next
to the resident FFT kernel's time per world-step.  One JSON line (committed under profiles/)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import lenia_oracle as lo  # noqa: E402  (tool, not product: the fp64 reference of the potential)

DEV = 'cuda:0'
n, B = 128, 2048
cfg, worlds = bench.make_worlds_numpy(B, 3)
K64 = lo.get_kernels_and_mapping([dict(p) for p in bench.ORBIUM_KP], [n, n], 1, 13, True, np.float64)[0][0, 0, 0]
ref = np.real(np.fft.ifft2(np.fft.fft2(worlds[:8, 0].astype(np.float64)) * K64))
jk = np.outer(np.arange(n), np.arange(n))
F = np.exp(-2j * np.pi * jk / n)
Fr, Fi = torch.tensor(F.real, dtype=torch.float32, device=DEV), torch.tensor(F.imag, dtype=torch.float32, device=DEV)
Kr, Ki = torch.tensor(K64.real / (n * n), dtype=torch.float32, device=DEV), torch.tensor(K64.imag / (n * n), dtype=torch.float32, device=DEV)
x = torch.from_numpy(worlds[:, 0]).to(DEV)


def split(a, mode):
    if mode == 'tf32x3':
        hi = (a.view(torch.int32) & ~0x1fff).view(torch.float32)  # 10 explicit mantissa bits
        return hi, a - hi
    hi = a.to(torch.bfloat16).float()
    return hi, a - hi


def mm(a, b, mode):
    if mode in ('fp32', 'tf32'):
        return a @ b
    (ah, al), (bh, bl) = split(a, mode), split(b, mode)
    if mode == 'bf16x3':
        f = lambda p, q: (p.to(torch.bfloat16) @ q.to(torch.bfloat16)).float()  # noqa: E731
        return f(ah, bh) + f(ah, bl) + f(al, bh)
    return ah @ bh + ah @ bl + al @ bh


def potential(x, mode):
    # rows: Y = x F (complex), columns: Z = F Y; multiply; inverse with conj(F)
    yr, yi = mm(x, Fr, mode), mm(x, Fi, mode)
    zr, zi = mm(Fr, yr, mode) - mm(Fi, yi, mode), mm(Fr, yi, mode) + mm(Fi, yr, mode)
    pr, pi = zr * Kr - zi * Ki, zr * Ki + zi * Kr
    ur, ui = mm(pr, Fr, mode) + mm(pi, Fi, mode), mm(pi, Fr, mode) - mm(pr, Fi, mode)  # times conj(F) from the right
    return mm(Fr, ur, mode) + mm(Fi, ui, mode)  # real part of conj(F) U


out = {}
for mode in ('fp32', 'tf32', 'tf32x3', 'bf16x3'):
    torch.backends.cuda.matmul.allow_tf32 = mode in ('tf32', 'tf32x3')
    p = potential(x, mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        p = potential(x, mode)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    out[mode] = {'linf_vs_fp64': float(np.abs(p[:8].cpu().numpy() - ref).max()), 'us_per_world_potential': 1e6 * dt / B}
torch.backends.cuda.matmul.allow_tf32 = False
out['note'] = ('resident FFT kernel: potential within 3e-7 of fp64, 0.04 us of GPU time per world-step INCLUDING growth, update and statistics '
               '(163 ms / (4096 x 1024)); the GEMM form is the forward and inverse transforms only')
print(json.dumps(out))
