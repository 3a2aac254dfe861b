#!/usr/bin/env python
"""Opcode histogram of the built library (cuobjdump -sass), per kernel family: the evidence that the resident kernels use tensor memory
(LDTM / STTM) and packed FP32 (FADD2 / FMUL2 / FFMA2), that no tensor-core (UTC*MMA / HMMA) or TMA (UTMA*) instruction is involved (an FFT
is not a GEMM, and thread-private 4-byte columns cannot be bulk-stored), and that cp.async (LDGSTS) stages the multipliers of gen2.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'leniax_b200', 'libleniax_b200.so')
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
fams = collections.OrderedDict((k, collections.Counter()) for k in ('lnx_world128_tm', 'lnx_world128_gen2', 'lnx_world128_gen_tm', 'lnx_world128_generic',
                                                                  't64h::', 't64::', 't2k::', 'tiled::', 'setup::', 'other'))
n_kernels = collections.Counter()
fam = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        fam = next((k for k in fams if k in name), 'other')
        n_kernels[fam] += 1
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and fam:
        fams[fam][m.group(1)] += 1
keys = ['LDTM', 'STTM', 'FADD2', 'FMUL2', 'FFMA2', 'FADD', 'FMUL', 'FFMA', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'ACQBULK', 'PREEXIT', 'UTMALDG', 'UTMASTG', 'UTCHMMA', 'UTCQMMA', 'HMMA']
print('%-22s %8s %9s ' % ('kernel family', 'kernels', 'instr') + ' '.join('%7s' % k for k in keys))
for k, c in fams.items():
    if n_kernels[k]:
        print('%-22s %8d %9d ' % (k, n_kernels[k], sum(c.values())) + ' '.join('%7d' % c.get(x, 0) for x in keys))
