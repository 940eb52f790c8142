"""ctypes harness around the REAL reference runtime compiled into oracle/_ref (see oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  The reference keeps its sample-rate tables and all dither/PRNG state in
process globals (runtime/dsp_runtime.c:36-38,103-110; runtime/dsp_tpdf.h:11-13,23,33), so one loaded
library == one live program.  We therefore run streams one after another ("stream-major"): private
[code|data] buffer, dspRuntimeInit(seed), then all frames, cores ascending inside each frame with one
io[] shared by all cores (canonical order).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import wire

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
IOMAX = 32


def available(fmt: int = 2, strict: bool = True) -> bool:
    return os.path.exists(_path(fmt, strict))


def _path(fmt, strict):
    return os.path.join(REFDIR, f"libavdspruntime{fmt}{'_strict' if strict else ''}.so")


_libs = {}


def lib(fmt: int, strict: bool = True):
    key = (fmt, strict)
    if key not in _libs:
        L = C.CDLL(_path(fmt, strict))
        L.dspRuntimeInit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.dspRuntimeInit.restype = C.c_int
        L.dspRuntimeReset.argtypes = [C.c_int, C.c_int, C.c_int]
        L.dspRuntimeReset.restype = C.c_int
        L.dspFindCore.argtypes = [C.c_void_p, C.c_int]
        L.dspFindCore.restype = C.c_void_p
        L.dspFindCoreBegin.argtypes = [C.c_void_p]
        L.dspFindCoreBegin.restype = C.c_void_p
        fn = getattr(L, f"dspRuntime_{fmt}")
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        _libs[key] = L
    return _libs[key]


class RefProgram:
    """One loaded program in the reference runtime (process-global: only one alive per (fmt,strict) lib)."""

    def __init__(self, words, fmt, fs, seed=0, dither=31, strict=True, max_words=None):
        self.L = lib(fmt, strict)
        self.fmt = fmt
        words = np.asarray(words, dtype=np.int32)
        h = wire.header(words)
        n = h["totalLength"] + h["dataSize"]
        self.size = max_words or (n + 16)
        self.buf = np.zeros(max(self.size, len(words)) + 2, dtype=np.int32)
        self.buf[: len(words)] = words
        self.rc = self.L.dspRuntimeInit(self.buf.ctypes.data, self.size, fs, seed, dither)
        if self.rc < 0:
            return
        self.total = self.rc
        self.data_ptr = self.buf.ctypes.data + 4 * self.total
        self.cores = []
        for k in range(1, 33):
            p = self.L.dspFindCore(self.buf.ctypes.data, k)
            if not p:
                break
            self.cores.append(self.L.dspFindCoreBegin(p))
            if h["numCores"] <= 1 and k == 1 and p == self.buf.ctypes.data:
                break   # no DSP_CORE in the program: dspFindCore returns the header for any k (App. C #8)
        self.run = getattr(self.L, f"dspRuntime_{fmt}")
        self.ins, self.outs = wire.io_maps(words)

    @property
    def data(self):
        return self.buf[self.total: self.total + wire.header(self.buf)["dataSize"]]

    def process(self, x: np.ndarray, in_idx=None, out_idx=None) -> np.ndarray:
        """x: [T, nIn] int32 (float bit patterns for fmt 5/6) -> [T, nOut] int32, canonical order."""
        in_idx = self.ins if in_idx is None else in_idx
        out_idx = self.outs if out_idx is None else out_idx
        x = np.ascontiguousarray(x, dtype=np.int32)
        T = x.shape[0]
        y = np.zeros((T, len(out_idx)), dtype=np.int32)
        io = np.zeros(IOMAX, dtype=np.int32)
        iop = io.ctypes.data
        run, cores, dp = self.run, self.cores, self.data_ptr
        for n in range(T):
            io[:] = 0
            io[in_idx] = x[n]
            for c in cores:
                run(c, dp, iop)
            y[n] = io[out_idx]
        return y
