"""The reference runtime's own entry points (runtime/dsp_runtime.h:160-164) as exported by
libavdsp_b200.so, wrapped the way a C host uses them: one caller-owned [code | data] buffer, one live
program per process, one core and one frame per dspRuntime_<fmt> call."""
from __future__ import annotations

import numpy as np

from . import _lib


class RuntimeCompat:
    def __init__(self, words, fmt: int, max_words=None):
        self.L = _lib.lib()
        self.fmt = fmt
        words = np.asarray(words, dtype=np.int32)
        total, dsize = int(words[1]), int(words[2])
        self.size = max_words if max_words is not None else total + dsize + 16
        self.buf = np.zeros(max(self.size, len(words)) + 2, dtype=np.int32)
        self.buf[: len(words)] = words
        self.run = getattr(self.L, f"dspRuntime_{fmt}")

    def init(self, fs: int, seed: int = 0, dither: int = 31) -> int:
        self.rc = self.L.dspRuntimeInit(self.buf.ctypes.data, self.size, fs, seed, dither)
        if self.rc > 0:
            self.total = self.rc
            self.cores = []
            for k in range(1, 33):
                p = self.L.dspFindCore(self.buf.ctypes.data, k)
                if not p:
                    break
                self.cores.append(self.L.dspFindCoreBegin(p))
                if p == self.buf.ctypes.data:
                    break
        return self.rc

    def reset(self, fs: int, seed: int = 0, dither: int = 31) -> int:
        return self.L.dspRuntimeReset(fs, seed, dither)

    def frame(self, io: np.ndarray) -> int:
        """cores ascending on one shared io[32] (canonical order)."""
        dp = self.buf.ctypes.data + 4 * self.total
        for c in self.cores:
            rc = self.run(c, dp, io.ctypes.data)
            if rc < 0:
                return rc
        return 0
