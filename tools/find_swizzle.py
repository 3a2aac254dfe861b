#!/usr/bin/env python
"""Builds the E2' unit permutation table of the R16 kernel (leniax_b200/csrc/lnx_w128r.cuh: E2_SWZ_LO / E2_SWZ_HI).

Lanes (a, h) of a quarter-warp (a = 4q..4q+3, h = 0/1) store / load 16-byte units of the spectral columns
k1' = a (h = 0) or 32 - a (16 for a = 0) (h = 1), and of those columns + 32.  The bank group of a unit is
4 * (col & 1) + (unit ^ g(col)); the eight lanes of a quarter-warp must hit eight different groups, i.e. g must take
four different values on the even columns and on the odd columns of every quarter.  Those column sets are disjoint,
so g(col) = rank of col inside its set.  tests/test_emulator.py re-checks every access pattern from the real address
functions."""


def k1_of(a, h):
    return a if h == 0 else (16 if a == 0 else 32 - a)


g = {}
for q in range(4):
    for base in (0, 32):
        cols = [k1_of(a, h) + base for a in range(4 * q, 4 * q + 4) for h in (0, 1)]
        for parity in (0, 1):
            for rank, c in enumerate(sorted(c for c in cols if c % 2 == parity)):
                assert c not in g
                g[c] = rank
assert sorted(g) == list(range(64))
lo = sum(g[c] << (2 * c) for c in range(32))
hi = sum(g[c + 32] << (2 * c) for c in range(32))
print(f'constexpr unsigned long long E2_SWZ_LO = {lo:#018x}ULL, E2_SWZ_HI = {hi:#018x}ULL;')
