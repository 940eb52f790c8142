"""ctypes wrapper of the CPU restatement oracle (oracle/avdsp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import wire

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_avdsp.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, ci = C.c_void_p, C.c_int
        L.avo_create.argtypes = [C.POINTER(vp), vp, ci, ci, ci, ci, ci, ci]
        L.avo_create.restype = ci
        L.avo_destroy.argtypes = [vp]
        L.avo_reset.argtypes = [vp, ci, ci, ci]
        L.avo_num_cores.argtypes = [vp]
        L.avo_total_length.argtypes = [vp]
        L.avo_data_size.argtypes = [vp]
        L.avo_code.argtypes = [vp]; L.avo_code.restype = vp
        L.avo_data.argtypes = [vp]; L.avo_data.restype = vp
        L.avo_get_aux.argtypes = [vp, vp]
        L.avo_set_aux.argtypes = [vp, vp]
        L.avo_run_core.argtypes = [vp, ci, vp]
        L.avo_run_frame.argtypes = [vp, vp]
        L.avo_process.argtypes = [vp, vp, vp, ci, vp, ci, vp, ci]
        L.avo_process_plugin_order.argtypes = [vp, vp, vp, ci, ci, ci, ci]
        _lib = L
    return _lib


class Oracle:
    """One stream instance of the oracle."""

    def __init__(self, words, fmt, fs, seed=0, dither=31, max_words=None):
        L = lib()
        self.L = L
        words = np.ascontiguousarray(words, dtype=np.int32)
        self.words = words
        h = C.c_void_p()
        mw = max_words if max_words is not None else (1 << 30)
        self.rc = L.avo_create(C.byref(h), words.ctypes.data, len(words), mw, fmt, fs, seed, dither)
        self.h = h if self.rc > 0 else None
        if self.h:
            self.ins, self.outs = wire.io_maps(words)
            self.total = L.avo_total_length(self.h)
            self.dsize = L.avo_data_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.avo_destroy(self.h)
            self.h = None

    @property
    def data(self) -> np.ndarray:
        p = self.L.avo_data(self.h)
        return np.ctypeslib.as_array((C.c_int32 * self.dsize).from_address(p)) if self.dsize else np.zeros(0, np.int32)

    @property
    def code(self) -> np.ndarray:
        p = self.L.avo_code(self.h)
        return np.ctypeslib.as_array((C.c_int32 * self.total).from_address(p))

    def aux(self) -> np.ndarray:
        a = np.zeros(8, np.int32)
        self.L.avo_get_aux(self.h, a.ctypes.data)
        return a

    def process(self, x, in_idx=None, out_idx=None) -> np.ndarray:
        in_idx = np.asarray(self.ins if in_idx is None else in_idx, dtype=np.int32)
        out_idx = np.asarray(self.outs if out_idx is None else out_idx, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.int32).reshape(-1, max(len(in_idx), 1))
        T = x.shape[0]
        y = np.zeros((T, len(out_idx)), dtype=np.int32)
        self.L.avo_process(self.h, x.ctypes.data, y.ctypes.data, T,
                           in_idx.ctypes.data, len(in_idx), out_idx.ctypes.data, len(out_idx))
        return y

    def process_plugin_order(self, x, period, n_in, n_out) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.int32).reshape(-1, n_in)
        y = np.zeros((x.shape[0], n_out), dtype=np.int32)
        self.L.avo_process_plugin_order(self.h, x.ctypes.data, y.ctypes.data, x.shape[0], period, n_in, n_out)
        return y


def run_streams(words, fmt, fs, x, seeds=None, dither=31):
    """x: [S, T, nIn] -> [S, T, nOut]; also returns the list of final (data, aux) per stream."""
    S = x.shape[0]
    outs, states = [], []
    for s in range(S):
        o = Oracle(words, fmt, fs, seed=int(seeds[s]) if seeds is not None else s, dither=dither)
        assert o.rc > 0, o.rc
        outs.append(o.process(x[s]))
        states.append((o.data.copy(), o.aux(), o.code.copy()))
    return np.stack(outs), states
