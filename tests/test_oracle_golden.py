"""Pins the CPU oracle (oracle/avdsp_oracle.c) against the golden vectors produced by the real
reference runtime (tests/golden/make_golden.py) and, when it was built here, against the compiled
reference itself (oracle/_ref)."""
import numpy as np
import pytest

from conftest import CASES, load_program, load_vector, vector_names
from avdsp_b200 import synth


@pytest.mark.parametrize("name", vector_names())
def test_oracle_matches_reference_golden(oracle_lib, name):
    v = load_vector(name)
    w = load_program(v["program"])
    o = oracle_lib.Oracle(w, v["fmt"], v["fs"], seed=v["seed"], dither=v["dither"])
    assert o.rc > 0
    y = o.process(v["x"])
    assert np.array_equal(y, v["y"]), f"{name}: oracle output differs from the reference's"
    assert np.array_equal(o.data, v["data"]), f"{name}: oracle data area differs from the reference's"
    assert np.array_equal(o.code, v["code"]), f"{name}: MEM words in the code area differ"


def test_golden_inputs_are_the_documented_synthetic_pcm():
    for name in vector_names():
        v = load_vector(name)
        gen = synth.pcm_float if v["fmt"] >= 5 else synth.pcm
        x = gen(v["stimulus"], 1, v["frames"], v["x"].shape[1], v["fs"])[0]
        assert np.array_equal(x, v["x"]), name


@pytest.mark.parametrize("prog,fmt,fs", CASES)
def test_oracle_matches_compiled_reference(oracle_lib, prog, fmt, fs):
    from oracle import refdriver
    if not refdriver.available(fmt):
        pytest.skip("oracle/_ref not built here (it needs /root/reference)")
    w = load_program(prog)
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    for kind, seed, dither in (("full", 9, 24), ("sine", 0, 31)):
        o = oracle_lib.Oracle(w, fmt, fs, seed=seed, dither=dither)
        x = gen(kind, 1, 700, len(o.ins), fs)[0]
        r = refdriver.RefProgram(w, fmt, fs, seed=seed, dither=dither)
        assert r.rc == o.rc > 0
        assert np.array_equal(r.process(x), o.process(x))
        assert np.array_equal(r.data, o.data)


def test_oracle_return_codes_match_reference(oracle_lib):
    """dspRuntimeInit / dspRuntimeReset negative codes (runtime/dsp_runtime.c:116-195)."""
    from oracle import refdriver
    w = load_program("c2_testrpi_xover_f2_192k")
    bad_sum = w.copy(); bad_sum[3] ^= 1
    bad_hdr = w.copy(); bad_hdr[0] = 0
    cases = [(w, 192000, None), (w, 48000, None), (w, 12345, None), (bad_sum, 192000, None),
             (bad_hdr, 192000, None), (w, 192000, 100)]
    expect = [len(w), -2, -1, -4, -1, -6]
    for (words, fs, mx), exp in zip(cases, expect):
        o = oracle_lib.Oracle(words, 2, fs, max_words=mx)
        assert o.rc == exp
        if refdriver.available(2):
            r = refdriver.RefProgram(words, 2, fs, max_words=mx)
            assert r.rc == exp


def test_oracle_chunking_is_invisible(oracle_lib):
    w = load_program("c1_crossover2x2lfe_f2_48k")
    x = synth.pcm("noise", 1, 600, 2, 48000)[0]
    a = oracle_lib.Oracle(w, 2, 48000).process(x)
    o = oracle_lib.Oracle(w, 2, 48000)
    b = np.concatenate([o.process(x[:1]), o.process(x[1:333]), o.process(x[333:])])
    assert np.array_equal(a, b)
