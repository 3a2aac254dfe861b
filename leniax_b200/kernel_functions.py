"""Radial kernel shell functions (setup side of the hot path, reference: leniax/kernel_functions.py:7-261).

Evaluated once per kernel when the kernels are rasterised; torch ops on whatever device the caller uses.
"""
from typing import Callable, Dict

import torch

from .constant import EPSILON


def poly_quad(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:7-37
    return (4 * X * (1 - X))**params[0]


def gauss_bump(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:40-69
    q = params[0]
    return torch.exp(q * (q - 1 / (X * (1 - X) + EPSILON)))


def step(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:72-102
    q = params[0]
    return ((X >= q) & (X <= 1 - q)).to(X.dtype)


def gauss(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:105-134
    q = params[0]
    return torch.exp(-(((X - q) / (0.3 * q))**2) / 2)


def threshold(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:137-166
    return (X >= params[0]).to(X.dtype)


def staircase(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:169-205
    m, s = params[0], params[1]
    out = 0.5 * ((X >= m - s) & (X < m - s / 2)).to(X.dtype)
    out = out + ((X >= m - s / 2) & (X <= m + s / 2)).to(X.dtype)
    return out + 0.5 * ((X > m + s / 2) & (X <= m + s)).to(X.dtype)


def triangle(params, X: torch.Tensor) -> torch.Tensor:  # kernel_functions.py:208-250
    m, s = params[0], params[1]
    left, right = m - s, m + s
    out = ((X >= left) & (X < m)).to(X.dtype) * (X - left) / (m - left)
    return out + ((X >= m) & (X <= right)).to(X.dtype) * (X - right) / (m - right)


register: Dict[str, Callable] = {
    'poly_quad': poly_quad,
    'gauss_bump': gauss_bump,
    'gauss': gauss,
    'step': step,
    'threshold': threshold,
    'staircase': staircase,
    'triangle': triangle,
}
