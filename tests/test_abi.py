"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/avdsp_b200.h
declares, validates programs with the reference's return codes before touching CUDA, and refuses to
compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_program
from avdsp_b200 import _lib, AvdspError, Executor


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "avdsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(dsp[A-Z]\w*|avdsp_b200_[a-z_0-9]+)\s*[\(\[]", text))
    names.discard("avdsp_b200_t")
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    decl = declared_symbols()
    assert decl, "header parse failed"
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/avdsp_b200.h but not exported"
    assert decl == set(_lib.SYMBOLS), decl ^ set(_lib.SYMBOLS)


def test_built_for_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_validation_codes_come_before_cuda():
    w = load_program("c2_testrpi_xover_f2_192k")
    def rc(words, fs=192000, fmt=2):
        try:
            Executor(words, fs, fmt, 2).close()
            return 0
        except AvdspError as e:
            return e.code
    bad_sum = w.copy(); bad_sum[3] ^= 1
    bad_hdr = w.copy(); bad_hdr[0] = 0
    newer = w.copy(); newer[6] = (200 << 16) | 28
    assert rc(bad_hdr) == -1
    assert rc(w, fs=12345) == -1
    assert rc(w, fs=48000) == -2
    assert rc(bad_sum) == -4
    assert rc(newer) == -5
    assert rc(w[:100]) == -6
    assert rc(w, fmt=3) == -7          # Q4.28 program on a float runtime
    assert rc(w, fmt=9) == -7


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = load_program("c2_testrpi_xover_f2_192k")
    with pytest.raises(AvdspError) as ei:
        Executor(w, 192000, 2, 2)
    assert ei.value.code == -10 and "no CPU fallback" in str(ei.value)


def test_reference_helpers_on_host():
    """dspFindCore / dspFindCoreBegin / dspQM32 / dspOpcodeText need no GPU (runtime/dsp_runtime.c:42-77)."""
    import ctypes as C
    L = _lib.lib()
    w = load_program("c2_testrpi_xover_f2_192k")
    buf = np.ascontiguousarray(w)
    base = buf.ctypes.data
    cores = [L.dspFindCore(base, k) for k in (1, 2, 3, 4)]
    assert [(p - base) // 4 if p else None for p in cores] == [12, 74, 287, None]
    begins = [(L.dspFindCoreBegin(p) - base) // 4 for p in cores[:3]]
    assert begins == [51, 245, 290]
    assert L.dspQM32(0.5, 28) == 1 << 27 and L.dspQM32(8.0, 28) == 0x7FFFFFFF and L.dspQM32(-9.0, 28) == -(1 << 31)
    assert L.dspQM64(0.25, 40) == 1 << 38 and L.dspQNM(1.0, 4, 28) == 1 << 28
    txt = (C.c_char_p * 62).in_dll(L, "dspOpcodeText")
    assert txt[0] == b"DSP_END_OF_CODE" and txt[50] == b"DSP_BIQUADS" and txt[61] == b"DSP_SINE"


def test_program_file_readers():
    from avdsp_b200 import program
    h = load_program("ref_dac8prodsp")            # .hex C-array form written by dspcreate -hexfile
    assert program.header(h)["totalLength"] == len(h) or program.header(h)["totalLength"] <= len(h)
    b = load_program("ref_dacdiy1")
    assert program.header(b)["numCores"] == 4 and program.header(b)["encoding"] == 28
