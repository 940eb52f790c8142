"""k_dag (kernel_dag.cu): programs that route signals through the X/Y registers -- subtractive crossovers, forks, sums of MEM
words -- as a DAG of cascades.  The reference's own fixtures of that kind (osx/crossoverLV6.bin, dacfabriceo*.bin,
windows/mydspcode.bin) and BASELINE config C1 run on it; bit-exact outputs and state against the oracle."""
import numpy as np
import pytest

from conftest import load_program
from avdsp_b200 import Executor, AvdspError, synth, describe, KERNEL_GENERIC, KERNEL_AUTO, KERNEL_DAG
from oracle import wire

pytestmark = pytest.mark.gpu

DAG_CASES = [
    ("c1_crossover2x2lfe_f2_48k", 48000, KERNEL_AUTO),
    ("ref_crossoverLV6", 48000, KERNEL_AUTO), ("ref_crossoverLV6", 96000, KERNEL_AUTO),
    ("ref_dacfabriceo", 48000, KERNEL_AUTO), ("ref_dacfabriceo", 96000, KERNEL_AUTO),
    ("ref_dacfabriceo_oppo", 88200, KERNEL_AUTO),
    ("ref_lxmini_lv8", 192000, KERNEL_AUTO), ("ref_lxmini_lv8", 44100, KERNEL_AUTO),
    ("ref_win_mydspcode", 176400, KERNEL_AUTO),
    # programs the chain kernels take first: the DAG kernel on request
    ("c2_testrpi_xover_f2_192k", 192000, KERNEL_DAG), ("ref_lxmini_lr2", 96000, KERNEL_DAG), ("c3_peq16_f2_48k", 48000, KERNEL_DAG),
]


def expected_state(ex, st):
    data, aux, code = st
    blk = np.zeros(ex.state_words, dtype=np.int32)
    blk[: ex.data_size] = data
    blk[ex.aux_offset: ex.aux_offset + 7] = aux[:7]
    for k, wd in enumerate(ex.mem_words):
        blk[ex.mem_offset + 2 * k: ex.mem_offset + 2 * k + 2] = code[wd: wd + 2]
    return blk


@pytest.mark.parametrize("prog,fs,kernel", DAG_CASES)
@pytest.mark.parametrize("stim", ["full", "noise"])
def test_dag_kernel_bit_exact(oracle_lib, prog, fs, kernel, stim):
    w = load_program(prog)
    S, T = 37, 333                                   # ragged: not a multiple of the tile or of the warp
    seeds = np.arange(S, dtype=np.int32) * 3 + 1
    ex = Executor(w, fs, 2, S, seeds=seeds, dither=24)
    ex.set_kernel(kernel)
    x = synth.pcm(stim, S, T, ex.n_in, fs)
    ys, sts = oracle_lib.run_streams(w, 2, fs, x, seeds=seeds, dither=24)
    y = ex.process(x)
    assert ex.last_kernel == "dag"
    bad = np.count_nonzero(y != ys)
    assert bad == 0, f"{prog}: {bad}/{y.size} samples differ; outputs {sorted(set(np.nonzero(y != ys)[2].tolist()))}"
    for s in (0, 1, 17, S - 1):
        got, exp = ex.get_state(s), expected_state(ex, sts[s])
        assert np.array_equal(got, exp), f"{prog}: state of stream {s} differs at words {np.nonzero(got != exp)[0][:12]}"


@pytest.mark.parametrize("prog,fs", [("ref_dacfabriceo", 96000), ("c1_crossover2x2lfe_f2_48k", 48000), ("ref_crossoverLV6", 96000), ("ref_lxmini_lv8", 96000)])
def test_dag_any_split_into_calls_and_kernel_switches(oracle_lib, prog, fs):
    w = load_program(prog)
    S, T = 5, 700
    seeds = np.arange(S, dtype=np.int32) + 11
    x = synth.pcm("full", S, T, 0, fs) if False else None
    ex = Executor(w, fs, 2, S, seeds=seeds, dither=31)
    x = synth.pcm("full", S, T, ex.n_in, fs)
    ys, sts = oracle_lib.run_streams(w, 2, fs, x, seeds=seeds, dither=31)
    cuts = [0, 1, 2, 33, 64, 65, 97, 300, 301, 640, T]
    outs = []
    for k, (a, b) in enumerate(zip(cuts, cuts[1:])):
        ex.set_kernel(KERNEL_GENERIC if k % 3 == 2 else KERNEL_AUTO)       # every third piece on the interpreter: same state layout
        outs.append(ex.process(np.ascontiguousarray(x[:, a:b])))
    assert np.array_equal(np.concatenate(outs, axis=1), ys)
    for s in range(S):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s


def _delays_everywhere(us_raw, us_dp, us_post, fs=48000):
    """a subtractive crossover with all three delay kinds on PARAM words: DSP_DELAY on the raw sample, DSP_DELAY_DP on a cascade's
    accumulator handed over through a MEM word, DSP_DELAY behind a saturation with stores on both sides of it"""
    a = wire.Asm(fmt=2, fmin=fs, fmax=fs)
    a.core()
    a.tpdf_calc(24)
    a.param()
    lp = a.biquad_sections([[wire.rbj_peak(fs, 900, 0.8, 1.5)], [wire.rbj_peak(fs, 300, 1.1, 0.7)]])
    eq = a.biquad_sections([[wire.rbj_peak(fs, 2500, 2.0, 1.2)]])
    d_raw = a.delay_param(3000, us_raw, fs)
    d_dp = a.delay_param(3000, us_dp, fs)
    d_post = a.delay_param(3000, us_post, fs)
    m = a.mem_location()
    # producer: io 8 -> eq -> MEM
    a.load_gain(8, 0.7); a.biquads(eq); a.store_mem(m)
    # raw-sample delay path (crossoverLV6.c): delayed input minus the low-passed input
    a.load(9); a.simple("COPYXY"); a.delay(d_raw); a.gain(1.0); a.simple("SWAPXY"); a.gain(0.9); a.biquads(lp); a.simple("SUBYX")
    a.sat0db_tpdf(); a.store(0); a.delay(d_post); a.store(1)
    a.simple("SWAPXY"); a.sat0db_tpdf(); a.store(2)
    a.core()
    # accumulator delay path (oktodac_fabriceo.c): delayed MEM minus a cascade on the MEM, then >> 28, gain, another cascade
    a.load_mem(m); a.simple("COPYXY"); a.delay(d_dp, dp=True); a.simple("SWAPXY"); a.biquads(lp); a.simple("SUBYX")
    a.sat0db_gain(0.8); a.store(3)
    a.simple("SWAPXY"); a.shift(-100); a.gain(0.6); a.biquads(eq); a.sat0db_tpdf_gain(0.9); a.store(4)
    return a.end()


def test_dag_delay_lines_patched_mid_stream(oracle_lib):
    """reload_params shortens / lengthens every kind of delay between calls: stale ring indices (dsp_runtime.c:769-824 use a
    stale index once, then restart at 0) must be honoured by the DAG kernel exactly like by the interpreter"""
    fs, S = 48000, 4
    progs = [_delays_everywhere(2000, 1500, 1800), _delays_everywhere(300, 2500, 200), _delays_everywhere(2900, 100, 2950), _delays_everywhere(40, 60, 20)]
    assert "DAG kernel geometry" in describe(progs[0], fs, 2)
    seeds = np.arange(S, dtype=np.int32)
    ex = Executor(progs[0], fs, 2, S, seeds=seeds, dither=24)
    gen = Executor(progs[0], fs, 2, S, seeds=seeds, dither=24)
    gen.set_kernel(KERNEL_GENERIC)
    ora = [oracle_lib.Oracle(progs[0], 2, fs, seed=int(s), dither=24) for s in seeds]
    lens = [257, 41, 500, 96]
    for k, (p, n) in enumerate(zip(progs, lens)):
        if k:
            ex.reload_params(p); gen.reload_params(p)
            for o in ora:
                o.code[: len(p)] = p[: len(o.code)]
        x = synth.pcm("full", S, n, ex.n_in, fs)
        x = np.roll(x, 7 * k, axis=1)
        y, yg = ex.process(x), gen.process(x)
        assert ex.last_kernel == "dag" and gen.last_kernel == "generic"
        for s in range(S):
            assert np.array_equal(y[s], ora[s].process(x[s])), (k, s)
        assert np.array_equal(y, yg), k
        for s in range(S):
            assert np.array_equal(ex.get_state(s), gen.get_state(s)), (k, s)


def test_dag_full_width_batch(oracle_lib):
    """4096 distinct streams of the reference's dacfabriceo.bin (several CTA waves, partial last CTA)"""
    w, fs, S, T = load_program("ref_dacfabriceo"), 96000, 4096, 512
    seeds = np.arange(S, dtype=np.int32)
    ex = Executor(w, fs, 2, S, seeds=seeds, dither=24)
    x = synth.pcm("noise", S, T, ex.n_in, fs)
    y = ex.process(x)
    assert ex.last_kernel == "dag"
    for s in (0, 1, 23, 24, 25, 2000, S - 2, S - 1):
        o = oracle_lib.Oracle(w, 2, fs, seed=s, dither=24)
        assert np.array_equal(y[s], o.process(x[s])), s
        assert np.array_equal(ex.get_state(s)[: ex.data_size], o.data), s
