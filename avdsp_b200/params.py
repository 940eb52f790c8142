"""The `-dumpfile` symbol table of the reference's dspcreate (encoder/dsp_encoder.c:476-503, dsp_dump): one line per
parameter a host may rewrite in the loaded program,  `name offset paramNum size`  -- offset is relative to the first data word
of the DSP_PARAM_NUM section numbered paramNum, or the absolute word index when paramNum is 0; size in words.

Host-side helper for avdsp_b200_set_param / avdsp_b200_reload_params (the live parameter path, per stream or for all)."""
from __future__ import annotations

from typing import Dict, NamedTuple

import numpy as np


class Symbol(NamedTuple):
    name: str
    offset: int
    param_num: int
    size: int


def parse_dump(text: str) -> Dict[str, Symbol]:
    table = {}
    for ln in text.splitlines():
        f = ln.split()
        if len(f) != 4:
            continue
        try:
            table[f[0]] = Symbol(f[0], int(f[1]), int(f[2]), int(f[3]))
        except ValueError:
            continue
    return table


def load_dump(path) -> Dict[str, Symbol]:
    return parse_dump(open(path).read())


def word_index(words, sym: Symbol) -> int:
    """The same resolution as avdsp_b200_param_index, without an instance (walks the opcode chain of `words`)."""
    w = np.asarray(words).view(np.uint32)
    if sym.param_num == 0:
        return sym.offset
    p = 0
    while p < len(w):
        op, skip = int(w[p]) >> 16, int(w[p]) & 0xFFFF
        if skip == 0:
            break
        if op == 5 and int(np.int32(w[p + 1])) == sym.param_num:        # DSP_PARAM_NUM
            if not 0 <= sym.offset < skip - 2:
                raise ValueError(f"{sym.name}: offset outside its PARAM_NUM section")
            return p + 2 + sym.offset
        p += skip
    raise KeyError(f"{sym.name}: no DSP_PARAM_NUM {sym.param_num} in the program")
