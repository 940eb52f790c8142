"""Host logic of the kernel planners, through the C ABI's host-only entry point avdsp_b200_describe (no CUDA device needed):
which kernels can take a program and how k_chain3 cuts cascades into part warps and places them on the SM sub-partitions."""
import re

import pytest

from conftest import load_program
import avdsp_b200


def _v3(trace):
    for ln in trace.splitlines():
        if ln.startswith("chain kernel v3 geometry"):
            return ln
    return None


def test_c2_is_cut_into_fourteen_part_warps_with_flat_sub_partitions():
    w = load_program("c2_testrpi_xover_f2_192k")
    ln = _v3(avdsp_b200.describe(w, 192000, 2, n_streams=4096, num_sms=148))
    assert ln is not None
    assert "28 streams/CTA" in ln and "14 cascade warps of <= 4 sections" in ln
    parts = [tuple(map(int, m)) for m in re.findall(r" (\d+):(\d+)\+(\d+)@(\d+)", ln.split("parts (chain:first+n@base):")[1].split(";")[0])]
    per_chain = {}
    for chain, first, n, base in parts:
        per_chain.setdefault(chain, []).append((first, n, base))
    assert {c: sum(n for _, n, _ in v) for c, v in per_chain.items()} == {0: 4, 1: 4, 2: 6, 3: 8, 4: 6, 5: 6, 6: 8, 7: 6}
    for v in per_chain.values():
        v.sort()
        assert v[0][0] == 0 and v[0][2] == 0                       # a chain starts at its first section, on the frame being read
        for (f0, n0, b0), (f1, n1, b1) in zip(v, v[1:]):
            assert f1 == f0 + n0                                   # parts tile the cascade
            assert b1 == b0 + n0 - 1 + 32                          # the next part runs one tile behind the previous part's tail
    bins = ln.split("s = store warp):")[1].split("|")
    assert len(bins) == 4
    sums = sorted(sum(int(t) for t in b.split() if t.isdigit()) for b in bins)
    assert sums == [11, 11, 13, 13]                                # (4T,4,3,helper) x 2 + (4,3,3,3) x 2: see DESIGN.md 4.2a
    assert sorted(len(b.split()) for b in bins) == [4, 4, 4, 4]
    assert sum("d" in b.split() for b in bins) == 1 and sum("s" in b.split() for b in bins) == 1
    size = int(re.search(r"(\d+) B smem", ln).group(1))
    assert size <= 227 * 1024


@pytest.mark.parametrize("streams,per_cta", [(64, 1), (148 * 9, 9), (148 * 40, 32)])
def test_streams_per_cta_follow_the_sm_count(streams, per_cta):
    w = load_program("c2_testrpi_xover_f2_192k")
    ln = _v3(avdsp_b200.describe(w, 192000, 2, n_streams=streams, num_sms=148))
    assert ln is not None and f"{per_cta} streams/CTA" in ln


def test_programs_outside_the_v3_shape_are_left_to_the_other_kernels():
    for prog, fmt in (("c3_peq16_f2_48k", 2), ("c3_peq16_f3_48k", 3)):          # 16 sections, plain finish: four parts of four (fmt 2 and 3)
        t = avdsp_b200.describe(load_program(prog), 48000, fmt, n_streams=65536)
        assert "8 cascade warps of <= 4 sections" in _v3(t) and "0:12+4@105" in _v3(t) and "chain kernel v2 geometry" in t
    t = avdsp_b200.describe(load_program("c3_peq16_f4_48k"), 48000, 4, n_streams=65536)                # double ALU: no chain kernel
    assert _v3(t) is None
    t = avdsp_b200.describe(load_program("c5_mixer8x8_f2_192k"), 192000, 2, n_streams=4096)
    assert "time-parallel mix kernel: usable" in t and _v3(t) is None
    t = avdsp_b200.describe(load_program("c4_fir4096_f2_48k"), 48000, 2, n_streams=1024)
    assert "time-parallel FIR kernel: usable" in t


def test_describe_reports_the_reference_error_codes():
    w = load_program("c2_testrpi_xover_f2_192k").copy()
    w[3] ^= 1                                                       # checksum word
    with pytest.raises(avdsp_b200.AvdspError) as e:
        avdsp_b200.describe(w, 192000, 2)
    assert e.value.code == -4
    with pytest.raises(avdsp_b200.AvdspError) as e:
        avdsp_b200.describe(load_program("c2_testrpi_xover_f2_192k"), 44100, 2)    # fs outside the header range
    assert e.value.code == -2


def test_odd_cascade_lengths_are_cut_into_near_equal_parts():
    """Cascades of 1, 3, 5, 7 and 3 sections: 5 -> 3 + 2, 7 -> 4 + 3, the rest stay whole (the program of tests/test_gpu_chain3.py)."""
    from oracle import wire
    fs = 48000
    a = wire.Asm(fmt=2, fmin=fs, fmax=fs)
    a.core(); a.tpdf_calc(20)
    a.param()
    secs = [wire.rbj_peak(fs, f, q, g) for f, q, g in ((120.0, 0.7, 2.0), (900.0, 1.2, 0.5), (2500.0, 3.0, 1.8), (5200.0, 0.9, 0.6),
                                                         (9000.0, 2.0, 1.5), (300.0, 0.5, 1.2), (14000.0, 1.0, 0.8))]
    hdr = {n: a.biquad_sections([[c] for c in secs[:n]]) for n in (1, 3, 5, 7)}
    a.load(8); a.biquads(hdr[1]); a.sat0db(); a.store(0)
    a.load_gain(9, 1.0); a.biquads(hdr[3]); a.sat0db_tpdf(); a.delay_fixed_us(150, fs); a.store(1)
    a.core()
    a.load_gain(8, 0.35); a.biquads(hdr[5]); a.sat0db(); a.delay_fixed_us(840, fs); a.store(2)
    a.load_gain(9, -1.7); a.biquads(hdr[7]); a.sat0db_tpdf(); a.store(3); a.store(7)
    a.load_gain(8, 1.0); a.biquads(hdr[3]); a.sat0db(); a.store(0); a.store(4); a.store(5); a.store(6)
    ln = _v3(avdsp_b200.describe(a.end(), fs, 2, n_streams=148 * 20))
    assert ln is not None and "20 streams/CTA" in ln
    parts = [tuple(map(int, m)) for m in re.findall(r" (\d+):(\d+)\+(\d+)@(\d+)", ln.split("parts (chain:first+n@base):")[1].split(";")[0])]
    sizes = {}
    for chain, first, n, base in parts:
        sizes.setdefault(chain, []).append(n)
    assert sorted(map(tuple, sizes.values())) == [(1,), (3,), (3,), (3, 2), (4, 3)]
    # the ring must hold two tiles, the largest lag (second part of the 7-section chain: 3 + 32 + 2) and the longest delay (40)
    assert "largest lag 37 frames" in ln and "row ring 256 steps" in ln


def test_loop_order_analysis():
    """The decoder proves when the ALSA plugin's core-major loop nest equals the canonical order (no MEM word, io slot or
    dither value crosses a core boundary); only then does a plugin-order request keep the fused kernels."""
    def order(prog, fs, fmt=2):
        return [ln for ln in avdsp_b200.describe(load_program(prog), fs, fmt).splitlines() if ln.startswith("loop order")][0]
    for prog, fs in (("c2_testrpi_xover_f2_192k", 192000), ("c3_peq16_f2_48k", 48000), ("c5_mixer8x8_f2_192k", 192000)):
        assert "plugin order == canonical order" in order(prog, fs), prog
    assert "MEM word is shared" in order("ref_dacdiy1", 192000)
    assert "dither value another core computes" in order("ref_crossoverLV6", 48000)
    from oracle import wire
    a = wire.Asm(fmt=2)
    a.core(); a.load(8); a.store(0)
    a.core(); a.load(0); a.store(1)                      # io hand-off: core 2 reads what core 1 stored
    assert "handed from one core to another" in [ln for ln in avdsp_b200.describe(a.end(), 48000, 2).splitlines() if ln.startswith("loop order")][0]
    a = wire.Asm(fmt=2)
    a.core(); a.load(8); a.store(0)
    a.core(); a.tpdf_calc(20); a.load(9); a.store(1)     # the dither table (STORE mask) changes in core 2: core 1's first period differs
    assert "switches the dither table" in [ln for ln in avdsp_b200.describe(a.end(), 48000, 2).splitlines() if ln.startswith("loop order")][0]


def test_dump_file_symbol_table_resolves_to_program_words():
    """`dspcreate -dumpfile` (encoder/dsp_encoder.c:476-503): name, offset, PARAM_NUM number, size"""
    import os
    from conftest import GOLDEN
    from avdsp_b200 import params
    from oracle import wire
    w = load_program("c1_crossover2x2lfe_f2_48k")
    t = params.load_dump(os.path.join(GOLDEN, "programs", "c1_crossover2x2lfe_f2_48k.dump"))
    assert t["BQ2_LOWPASS_1"] == params.Symbol("BQ2_LOWPASS_1", 0, 1, 28)
    for name in ("BQ2_LOWPASS_1", "BQ2_HIGHPASS_3", "BQ4_EQ_LFE_-1", "BQ6_PRE_FILTER"):
        i = params.word_index(w, t[name])
        assert (int(w[i]) >> 16) & 0xFFFF == wire.OP["BIQUADS"], name      # the section header the BIQUADS opcode points to
    i = params.word_index(w, t["DELAY_HIGH_LOW_1"])
    assert int(w[i]) & 0xFFFF == 294 and int(w[i]) >> 16 == 71             # microseconds | max samples << 16 (encoder :1111-1118)


def test_float_class_needs_coefficients_it_can_bound():
    """The float class of the chain kernels guards its hardware multiplies through their operands, which needs every non-zero
    biquad coefficient within [2^-60, 2^7) (avdsp_dev.cuh, fltGuard); anything else stays on the interpreter, and says so."""
    from oracle import wire
    fs = 48000

    def prog(b0):
        a = wire.Asm(fmt=3, fmin=fs, fmax=fs)
        a.core(); a.param()
        e = a.biquad_sections([[(b0, 0.1, 0.05, 0.3, -0.2)], [wire.rbj_peak(fs, 1000.0, 1.0, 1.5)]])
        a.load_gain(8, 0.5); a.biquads(e); a.sat0db(); a.store(0)
        return a.end()

    ok = avdsp_b200.describe(prog(1.5), fs, 3, 4096)
    assert "chain kernel v2 geometry" in ok and "chain kernels not used" not in ok
    for b0 in (200.0, 1e-20):
        t = avdsp_b200.describe(prog(b0), fs, 3, 4096)
        assert "chain kernel v2 geometry" not in t and "outside [2^-60, 2^7)" in t, t
    # the same coefficients in fixed point are no concern
    a = wire.Asm(fmt=2, fmin=fs, fmax=fs)
    a.core(); a.param()
    e = a.biquad_sections([[(7.5, 0.1, 0.05, 0.3, -0.2)]])
    a.load_gain(8, 0.5); a.biquads(e); a.sat0db(); a.store(0)
    assert "chain kernel v2 geometry" in avdsp_b200.describe(a.end(), fs, 2, 4096)
