"""Cell codecs needed to start a simulation from a leniax config (reference: leniax/loader.py:16-129, 206-350).

Only decoding (and the quantisation used by the initialisers) is provided: gzip+base64 int32, the 2-character code and
the legacy run-length format used by the reference's test fixtures.
"""
import base64
import gzip
import io
import pickle
import re
from typing import Dict, List

import numpy as np
import torch

from .constant import NB_CHARS

_MAX_VAL = NB_CHARS**2 - 1


def make_array_compressible(cells: torch.Tensor) -> torch.Tensor:  # loader.py:16-30
    # (a CUDA tensor divided by a Python scalar is multiplied by the scalar's reciprocal: 1 ulp off k / 12543 on 28 % of the cells,
    # measured in profiles/r2_parity.txt; tensor / tensor is the correctly rounded IEEE quotient XLA's divide gives)
    ints = torch.round(cells * _MAX_VAL).to(torch.int32)
    return ints.to(torch.float32) / torch.full((1, ), float(_MAX_VAL), dtype=torch.float32, device=cells.device)


def decompress_array_gzip(string_cells: str) -> torch.Tensor:  # loader.py:105-129
    ints = np.frombuffer(gzip.decompress(base64.b64decode(string_cells)), dtype='<i4')
    n = int(ints[0])
    shape = [int(s) for s in ints[1 + n:]]
    return torch.from_numpy((ints[1:1 + n].reshape(shape) / _MAX_VAL).astype(np.float32))


def ch2val(c: str) -> int:  # loader.py:148-175
    def idx(ch):
        if ord(ch) >= ord('À'):
            return ord(ch) - ord('À') + (ord('Z') - ord('A')) + (ord('z') - ord('a'))
        if ord(ch) >= ord('a'):
            return ord(ch) - ord('a') + (ord('Z') - ord('A'))
        return ord(ch) - ord('A')

    assert len(c) == 2
    return idx(c[0]) * NB_CHARS + idx(c[1])


def _legacy_val(ch: str) -> int:  # loader.py:275-283
    if ch in '.b':
        return 0
    if ch == 'o':
        return 255
    if len(ch) == 1:
        return ord(ch) - ord('A') + 1
    return (ord(ch[0]) - ord('p')) * 24 + (ord(ch[1]) - ord('A') + 25)


def deprecated_decompress_array(cells_code: str, nb_dims: int) -> torch.Tensor:
    """Legacy RLE: ``$`` ends a row, ``%`` a plane, ``#`` a volume; digits repeat; two-letter values start with p..y
    (loader.py:246-350)."""
    level = {'$': 1, '%': 2, '#': 3}
    closing = ['', '$', '%', '#'][nb_dims - 1]
    stacks: List[List] = [[] for _ in range(nb_dims)]
    prefix, count = '', ''
    for ch in cells_code.rstrip('!') + closing:
        if ch.isdigit():
            count += ch
        elif ch in 'pqrstuvwxy@':
            prefix = ch
        else:
            n = int(count) if count else 1
            tok = prefix + ch
            if tok in level:
                for d in range(level[tok]):
                    stacks[d + 1].append(stacks[d])
                    stacks[d + 1].extend([] for _ in range(n - 1))
                    stacks[d] = []
            else:
                stacks[0].extend([_legacy_val(tok) / 255] * n)
            prefix, count = '', ''
    nested = stacks[nb_dims - 1]
    lens = [0] * nb_dims

    def measure(d, lst):
        lens[d] = max(lens[d], len(lst))
        if d < nb_dims - 1:
            for sub in lst:
                measure(d + 1, sub)

    measure(0, nested)
    out = np.zeros(lens, dtype=np.float32)

    def fill(d, lst, idx):
        if d == nb_dims - 1:
            out[idx + (slice(0, len(lst)), )] = lst
        else:
            for i, sub in enumerate(lst):
                fill(d + 1, sub, idx + (i, ))

    fill(0, nested, ())
    return torch.from_numpy(out)


def compress_array(cells) -> str:  # loader.py:33-66
    """Cells state -> base64 string of the gzip-compressed int32 encoding (count, values, shape; little endian), the format
    ``decompress_array_gzip`` reads and ``LeniaIndividual.set_cells`` / ``set_init_cells`` store."""
    import codecs
    arr = cells.detach().cpu().numpy() if isinstance(cells, torch.Tensor) else np.asarray(cells)
    ints = np.round(arr.astype(np.float32) * np.float32(_MAX_VAL)).astype('<i4')
    shape_bytes = b''.join(int(d).to_bytes(4, 'little') for d in ints.shape)
    payload = int(ints.size).to_bytes(4, 'little') + ints.tobytes() + shape_bytes
    return str(codecs.encode(gzip.compress(payload), 'base64'), 'utf-8')


class _ArrayUnpickler(pickle.Unpickler):
    """Unpickles NumPy arrays only: the reference's plain-base64 format is a pickled array (loader.py:132-146) and configuration
    files are data, not code."""
    def find_class(self, module, name):
        if module.split('.')[0] == 'numpy' and name in ('_reconstruct', 'ndarray', 'dtype', 'scalar', '_frombuffer'):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'{module}.{name} is not allowed in a cells string')


def decompress_array_base64(string_cells: str) -> torch.Tensor:  # loader.py:132-146
    """base64 of a pickled array."""
    import codecs
    raw = codecs.decode(string_cells.encode(), 'base64')
    return torch.as_tensor(np.asarray(_ArrayUnpickler(io.BytesIO(raw)).load(), dtype=np.float32))


def decompress_array(string_cells: str, nb_dims: int = 0) -> torch.Tensor:  # loader.py:69-102 (best-effort chain, same order)
    import binascii
    try:
        return decompress_array_gzip(string_cells)
    except Exception:
        pass
    parts = string_cells.split('::')
    if len(parts) == 2 and len(parts[0]) % 2 == 0:
        try:
            shape = [int(c) for c in parts[1].split(';')]
            vals = [ch2val(parts[0][i:i + 2]) for i in range(0, len(parts[0]), 2)]
            ints = torch.tensor(vals, dtype=torch.int32).reshape(shape)
            return ints.to(torch.float32) / torch.tensor(float(_MAX_VAL))
        except Exception:
            pass
    try:  # an .npz archive stored as a latin1 string, uint8 cells under 'x'
        return torch.from_numpy((np.load(io.BytesIO(string_cells.encode('latin1')))['x'] / 255.).astype(np.float32))
    except Exception:
        pass
    try:
        return decompress_array_base64(string_cells)
    except (binascii.Error, pickle.UnpicklingError, ValueError, EOFError, IndexError, KeyError):
        pass
    # the legacy run-length format (loader.py:287-350): digits, '.', 'b', 'o', 'A'..'X', prefixes 'p'..'y' / '@', delimiters $ % #, final '!'
    if nb_dims >= 1 and re.fullmatch(r'[0-9.boA-Xp-y@$%#]*!?', string_cells.strip()):
        return deprecated_decompress_array(string_cells.strip(), nb_dims)
    raise ValueError(f'no decoder of leniax.loader matches this cells string ({len(string_cells)} characters)')


def load_raw_cells(config: Dict, use_init_cells: bool = True) -> torch.Tensor:  # loader.py:206-240
    nb_dims = config['world_params']['nb_dims']
    rp = config['run_params']
    cells = rp['init_cells'] if (use_init_cells and 'init_cells' in rp) else rp['cells']
    if isinstance(cells, str):
        if cells == 'MISSING':
            cells = torch.zeros(0)
        elif cells == 'last_frame.p':
            import os
            import pickle
            with open(os.path.join(config['main_path'], 'last_frame.p'), 'rb') as f:
                cells = torch.as_tensor(np.asarray(pickle.load(f), dtype=np.float32))
        else:
            cells = decompress_array(cells, nb_dims + 1)
    elif isinstance(cells, list):
        cells = torch.tensor(cells, dtype=torch.float32)
    if cells.dim() == nb_dims and config['world_params']['nb_channels'] == 1:
        cells = cells[None]
    return cells.to(torch.float32)
