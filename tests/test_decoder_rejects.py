"""Host decoder (avdsp_b200/csrc/decoder.cpp) on malformed / unsupported programs, through the host-only C-ABI call
avdsp_b200_describe: no CUDA device needed.  A crafted .bin must be refused at load time, never reach a kernel."""
import numpy as np
import pytest

from conftest import load_program
import avdsp_b200
from oracle import wire

ERR_MALFORMED, ERR_ENCODER_OLD = -11, -13
INT_MAX = 0x7FFFFFFF


def _refix(w):
    """recompute the header checksum after patching argument words (opcode words are what is summed)"""
    w = np.array(w, dtype=np.int32)
    s, _ = wire.checksum(w.view(np.uint32))
    w.view(np.uint32)[3] = s
    return w


def _code_of(words, fs=48000, fmt=2):
    with pytest.raises(avdsp_b200.AvdspError) as e:
        avdsp_b200.describe(words, fs, fmt)
    return e.value.code


def _find(w, name, nth=0):
    hits = [p for p, op, sk in wire.walk(w.view(np.uint32)) if op == wire.OP[name]]
    return hits[nth]


@pytest.mark.parametrize("name", ["old_rpi_dacfabriceo", "old_rpi_testrew", "old_osx_mydspcode"])
def test_encoder_0x100_files_are_rejected_by_name(name):
    """module_avdsp/rpi/*.bin: 11-word header, TPDF_CALC without its data word.  The reference runtime does not look at the
    version and crashes on rpi/dacfabriceo.bin; here: a named error."""
    w = load_program(name)
    fs = 48000
    assert _code_of(w, fs) == ERR_ENCODER_OLD
    assert "0x102" in avdsp_b200._lib.last_error()


def test_data_offsets_near_int_max_do_not_wrap():
    w = load_program("allops_gen_f2_multifs").copy()
    for op, argk in (("DELAY_1", 0), ("DITHER", 0), ("TPDF_CALC", 1), ("DCBLOCK", 0), ("DIRAC", 0), ("LOAD_MUX", None)):
        if argk is None:
            continue
        v = w.copy()
        v[_find(v, op) + 1 + argk] = INT_MAX - 1        # off + n would wrap negative in 32-bit arithmetic
        assert _code_of(_refix(v), 48000) == ERR_MALFORMED, op
    v = w.copy()
    v[_find(v, "RMS") + 2] = INT_MAX                    # delay * aluWords = 0x7FFFFFFF * 2 wraps to a small number in 32 bits
    assert _code_of(_refix(v), 48000) == ERR_MALFORMED
    v = w.copy()
    v[_find(v, "DISTRIB") + 2] = INT_MAX                # 1 + size
    assert _code_of(_refix(v), 48000) == ERR_MALFORMED
    v = w.copy()
    p = _find(v, "DELAY_DP")
    v[p + 1] = INT_MAX                                  # fixed delay in microseconds -> n * 2 words
    assert _code_of(_refix(v), 48000) == ERR_MALFORMED


@pytest.mark.parametrize("div", [-1, 64, 65, INT_MAX, -INT_MAX])
def test_data_table_divider_must_stay_inside_the_table(div):
    """index += div; if (index >= size) index -= size (runtime/dsp_runtime.c:911-912) walks out of the table otherwise"""
    w = load_program("allops_gen_f2_multifs").copy()
    p = _find(w, "DATA_TABLE")
    assert w[p + 3] == 64
    w[p + 2] = div
    assert _code_of(_refix(w), 48000) == ERR_MALFORMED
    w[p + 2] = 63
    assert "core 4" in avdsp_b200.describe(_refix(w), 48000, 2)      # the largest legal divider loads fine


def test_core_in_the_last_words_is_refused():
    a = wire.Asm(fmt=2)
    a.core(); a.load(8); a.store(0)
    w = a.end().copy()
    # turn the END_OF_CODE padding into a truncated DSP_CORE followed by nothing
    total = int(w[1])
    w[total - 2] = (wire.OP["CORE"] << 16) | 2
    w[total - 1] = 0
    w = w[:total]
    with pytest.raises(avdsp_b200.AvdspError):
        avdsp_b200.describe(_refix(w), 48000, 2)


def test_relative_pointers_outside_the_program():
    w = load_program("c2_testrpi_xover_f2_192k").copy()
    p = _find(w, "BIQUADS")
    w[p + 2] = INT_MAX
    assert _code_of(_refix(w), 192000) == ERR_MALFORMED
    w = load_program("c5_mixer8x8_f2_192k").copy()
    p = _find(w, "LOAD_MUX")
    w[p + 1] = -INT_MAX
    assert _code_of(_refix(w), 192000) == ERR_MALFORMED
