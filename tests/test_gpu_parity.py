"""Parity tests proper: the CUDA path, called through the C ABI (libavdsp_b200.so), against the CPU
oracle on the same seeded inputs, against the committed golden vectors of the real reference, and
through size-independent properties at larger sizes.

Bar: DSP_FORMAT 2 (int64 accumulator) is BIT-EXACT, outputs and every state word.  Formats 3..6 are
bit-exact in the generic executor as well (it restates the reference's IEEE helpers with integer ops), and so is the float
class of the chain kernels (hardware multiplies guarded by a per-stream exactness flag + interpreter re-execution); the one
kernel with a stated tolerance is the opt-in 3xTF32 tensor-core FIR.
"""
import numpy as np
import pytest

from conftest import CASES, load_program, load_vector, vector_names
from avdsp_b200 import (Executor, AvdspError, synth, INTERLEAVED, PLANAR, KERNEL_GENERIC, KERNEL_CHAIN, KERNEL_AUTO)

pytestmark = pytest.mark.gpu


def oracle_run(pyoracle, w, fmt, fs, x, seeds, dither):
    ys, sts = pyoracle.run_streams(w, fmt, fs, x, seeds=seeds, dither=dither)
    return ys, sts


def expected_state(ex, st):
    """Oracle (data, aux, code) -> the executor's per-stream state block."""
    data, aux, code = st
    blk = np.zeros(ex.state_words, dtype=np.int32)
    blk[: ex.data_size] = data
    blk[ex.aux_offset: ex.aux_offset + 7] = aux[:7]
    for k, wd in enumerate(ex.mem_words):
        blk[ex.mem_offset + 2 * k: ex.mem_offset + 2 * k + 2] = code[wd: wd + 2]
    return blk


def gen(fmt):
    return synth.pcm_float if fmt >= 5 else synth.pcm


@pytest.mark.parametrize("name", vector_names())
def test_golden_vectors_generic(name):
    v = load_vector(name)
    w = load_program(v["program"])
    ex = Executor(w, v["fs"], v["fmt"], 1, seeds=[v["seed"]], dither=v["dither"])
    ex.set_kernel(KERNEL_GENERIC)
    y = ex.process(v["x"][None])[0]
    assert np.array_equal(y, v["y"]), f"{name}: {np.count_nonzero(y != v['y'])} samples differ from the reference"
    st = ex.get_state(0)
    assert np.array_equal(st[: ex.data_size], v["data"])
    for k, wd in enumerate(ex.mem_words):
        assert np.array_equal(st[ex.mem_offset + 2 * k: ex.mem_offset + 2 * k + 2], v["code"][wd: wd + 2])


@pytest.mark.parametrize("name", [n for n in vector_names()])
def test_golden_vectors_auto_kernel(name):
    """Whatever kernel AUTO picks (the fused chain kernel where the program maps to it) must give the
    reference's bits, fixed point and float formats alike."""
    v = load_vector(name)
    w = load_program(v["program"])
    ex = Executor(w, v["fs"], v["fmt"], 1, seeds=[v["seed"]], dither=v["dither"])
    y = ex.process(v["x"][None])[0]
    assert np.array_equal(y, v["y"]), f"{name} ({ex.last_kernel}): {np.count_nonzero(y != v['y'])} samples differ"
    got, exp = ex.get_state(0)[: ex.data_size], v["data"]
    assert np.array_equal(got, exp), f"{name} ({ex.last_kernel}): state differs"


@pytest.mark.parametrize("prog,fmt,fs", CASES)
@pytest.mark.parametrize("kernel", [KERNEL_GENERIC, KERNEL_AUTO])
def test_parity_vs_oracle_multi_stream(oracle_lib, prog, fmt, fs, kernel):
    w = load_program(prog)
    S, T = 37, 300                      # ragged on purpose: not a multiple of any tile or warp size
    seeds = np.arange(S, dtype=np.int32) * 7 + 1
    ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
    ex.set_kernel(kernel)
    x = gen(fmt)("full" if fmt == 2 else "noise", S, T, ex.n_in, fs)
    ys, sts = oracle_run(oracle_lib, w, fmt, fs, x, seeds, 24)
    y = ex.process(x)
    bad = np.count_nonzero(y != ys)
    assert bad == 0, f"{prog} [{ex.last_kernel}]: {bad}/{y.size} samples differ"
    for s in (0, 1, S - 1):
        got, exp = ex.get_state(s), expected_state(ex, sts[s])
        diff = np.nonzero(got != exp)[0]
        # (the float class of the chain kernels included: streams that come near the underflow threshold or leave the binary32
        # range are re-executed by the interpreter, avdsp_dev.cuh fltGuard)
        assert diff.size == 0, f"{prog} [{ex.last_kernel}] stream {s}: state words {diff[:8]} differ"


@pytest.mark.parametrize("prog,fmt,fs", [c for c in CASES if c[1] == 2])
def test_chunked_calls_equal_one_call(prog, fmt, fs):
    """ALSA periods: any split of the frame range gives identical output and state."""
    w = load_program(prog)
    S, T = 9, 1500
    x = synth.pcm("noise", S, T, len(Executor(w, fs, fmt, 1).in_idx), fs)
    for kernel in (KERNEL_GENERIC, KERNEL_AUTO):
        a = Executor(w, fs, fmt, S); a.set_kernel(kernel)
        ya = a.process(x)
        b = Executor(w, fs, fmt, S); b.set_kernel(kernel)
        cuts = [0, 1, 2, 65, 66, 577, 1024, T]
        yb = np.concatenate([b.process(np.ascontiguousarray(x[:, c0:c1])) for c0, c1 in zip(cuts, cuts[1:])], axis=1)
        assert np.array_equal(ya, yb), (prog, a.last_kernel)
        for s in (0, S - 1):
            assert np.array_equal(a.get_state(s), b.get_state(s)), (prog, a.last_kernel)


def test_kernels_agree_and_can_alternate():
    """generic and chain kernels share one state layout: switching between them mid-stream is invisible."""
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 70, 900
    x = synth.pcm("full", S, T, 2, 192000)
    a = Executor(w, 192000, 2, S); a.set_kernel(KERNEL_GENERIC)
    ya = a.process(x)
    b = Executor(w, 192000, 2, S)
    parts = []
    for i, (c0, c1) in enumerate(((0, 100), (100, 433), (433, 434), (434, T))):
        b.set_kernel(KERNEL_CHAIN if i % 2 == 0 else KERNEL_GENERIC)
        parts.append(b.process(np.ascontiguousarray(x[:, c0:c1])))
    assert np.array_equal(ya, np.concatenate(parts, axis=1))
    assert np.array_equal(a.get_state(S - 1), b.get_state(S - 1))


def test_chain_kernel_is_selected_for_the_benchmark_programs():
    for prog, fs, kern in (("c2_testrpi_xover_f2_192k", 192000, "chain"), ("c5_mixer8x8_f2_192k", 192000, "mix"),
                           ("ref_dac8prodsp", 96000, "mix"),      # the DAC8PRO firmware program: raw LOAD_STORE copies + dithered gains
                           ("c3_peq16_f2_48k", 48000, "chain"), ("c3_peq16_f3_48k", 48000, "chain")):
        ex = Executor(load_program(prog), fs, 3 if "_f3_" in prog else 2, 64)
        ex.process(synth.pcm("noise", 64, 64, ex.n_in, fs))
        assert ex.last_kernel == kern, ex.trace
    for prog, fmt, fs, kern in (("c4_fir4096_f2_48k", 2, 48000, "fir"), ("c4_fir4096_f3_48k", 3, 48000, "fir"),
                                ("c4s_fir_f3_multifs", 3, 48000, "fir"), ("c4s_fir_f3_multifs", 3, 96000, "generic"),
                                ("c4s_fir_f4_multifs", 4, 48000, "generic")):
        ex = Executor(load_program(prog), fs, fmt, 8)
        ex.process(synth.pcm("noise", 8, 64, ex.n_in, fs))
        assert ex.last_kernel == kern, ex.trace
    ex = Executor(load_program("c1_crossover2x2lfe_f2_48k"), 48000, 2, 4)   # MEM hand-off, X/Y dataflow
    ex.process(synth.pcm("noise", 4, 64, ex.n_in, 48000))
    assert ex.last_kernel == "dag"                       # not a set of independent chains: a DAG of cascades (kernel_dag.cu)
    with pytest.raises(AvdspError):
        ex.set_kernel(KERNEL_CHAIN); ex.process(synth.pcm("noise", 4, 8, ex.n_in, 48000))


def test_planar_layout_and_device_path(oracle_lib):
    import torch
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 33, 257
    x = synth.pcm("noise", S, T, 2, 192000)
    ys, _ = oracle_run(oracle_lib, w, 2, 192000, x, np.zeros(S, np.int32), 31)
    for kernel in (KERNEL_GENERIC, KERNEL_AUTO):
        ex = Executor(w, 192000, 2, S); ex.set_kernel(kernel)
        yp = ex.process(np.ascontiguousarray(x.transpose(0, 2, 1)), layout=PLANAR)
        assert np.array_equal(yp.transpose(0, 2, 1), ys)
        ex2 = Executor(w, 192000, 2, S); ex2.set_kernel(kernel)
        yd = ex2.process(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(yd.cpu().numpy(), ys)


def _plugin_program():
    """ALSA convention (in io[8+k], out io[k]); core 2 dithers with the TPDF value core 1 computes, so the
    plugin's core-major order (core 2 sees the LAST value of the period) differs from the canonical order."""
    from oracle import wire
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.tpdf_calc(20); a.load_gain(8, 0.5); a.sat0db_tpdf(); a.store(0)
    a.core(); a.load_gain(9, 0.25); a.sat0db_tpdf(); a.store(1); a.load(8); a.store(2)
    return a.end()


def test_host_path_time_chunk_pipeline(oracle_lib, monkeypatch):
    """avdsp_b200_process with HOST buffers cuts the call into time chunks (copy-in / launch / copy-out pipeline):
    forcing small, ragged chunks must not change a bit, in both layouts."""
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 19, 701
    x = synth.pcm("full", S, T, 2, 192000)
    ys, _ = oracle_run(oracle_lib, w, 2, 192000, x, np.zeros(S, np.int32), 31)
    monkeypatch.setenv("AVDSP_B200_HOST_CHUNK", "96")
    ex = Executor(w, 192000, 2, S)
    assert np.array_equal(ex.process(x), ys)
    ex2 = Executor(w, 192000, 2, S)
    yp = ex2.process(np.ascontiguousarray(x.transpose(0, 2, 1)), layout=PLANAR)
    assert np.array_equal(yp.transpose(0, 2, 1), ys)


def test_alsa_sample_formats(oracle_lib):
    """S16_LE / S24_3LE input widened like linux/avdsp_plugin.c:109-121, S32 out."""
    from avdsp_b200.executor import PCM_S16, PCM_S24_3LE
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 3, 300
    x32 = synth.pcm("full", S, T, 2, 192000)
    x16 = (x32 >> 16).astype(np.int16)
    y16, _ = oracle_run(oracle_lib, w, 2, 192000, x16.astype(np.int32) << 16, np.zeros(S, np.int32), 31)
    assert np.array_equal(Executor(w, 192000, 2, S).process_pcm(x16, PCM_S16, T), y16)
    x24 = x32 & ~0xFF                                   # 24 significant bits
    b = np.empty(x24.shape + (3,), np.uint8)
    b[..., 0] = (x24 >> 8) & 0xFF; b[..., 1] = (x24 >> 16) & 0xFF; b[..., 2] = (x24 >> 24) & 0xFF
    y24, _ = oracle_run(oracle_lib, w, 2, 192000, x24, np.zeros(S, np.int32), 31)
    assert np.array_equal(Executor(w, 192000, 2, S).process_pcm(b, PCM_S24_3LE, T), y24)


def test_plugin_order_mode(oracle_lib):
    """core-major loop nest of linux/avdsp_plugin.c:95-142 with a given period."""
    fs, S, T, period = 48000, 3, 500, 128
    for w in (_plugin_program(), load_program("c2_testrpi_xover_f2_192k")):
        if w is not None and int(w[8]) == 9:
            fs = 192000
        ex = Executor(w, fs, 2, S, seeds=[0, 1, 2], dither=24)
        x = synth.pcm("noise", S, T, ex.n_in, fs)
        canon = Executor(w, fs, 2, S, seeds=[0, 1, 2], dither=24).process(x)
        ex.set_order(period)
        y = ex.process(x)
        # C2's cores do not talk to each other: the decoder proves both orders equal and keeps the fused kernel
        assert ex.last_kernel == ("chain" if int(w[8]) == 9 else "generic")
        nin, nout = max(ex.in_idx) - 8 + 1, max(ex.out_idx) + 1
        for s in range(S):
            o = oracle_lib.Oracle(w, 2, fs, seed=s, dither=24)
            xin = np.zeros((T, nin), np.int32)
            for k, slot in enumerate(ex.in_idx):
                xin[:, slot - 8] = x[s, :, k]
            yo = o.process_plugin_order(xin, period, nin, nout)
            assert np.array_equal(y[s], yo[:, ex.out_idx]), s
        if int(w[8]) != 9:
            assert not np.array_equal(y, canon), "the test program must distinguish the two orders"


def test_reference_entry_points_per_frame(oracle_lib):
    """dspRuntimeInit / dspFindCore / dspRuntime_2 exactly as a reference host drives them."""
    from avdsp_b200.compat import RuntimeCompat
    w = load_program("ref_dacdiy1")
    rt = RuntimeCompat(w, 2)
    assert rt.init(96000, seed=4, dither=24) == int(w[1])
    assert len(rt.cores) == 4
    o = oracle_lib.Oracle(w, 2, 96000, seed=4, dither=24)
    x = synth.pcm("noise", 1, 40, len(o.ins), 96000)[0]
    io = np.zeros(32, np.int32)
    oio = np.zeros(32, np.int32)
    for n in range(40):
        io[:] = 0; io[o.ins] = x[n]
        oio[:] = io
        assert rt.frame(io) == 0
        o.L.avo_run_frame(o.h, oio.ctypes.data)
        assert np.array_equal(io, oio), n
    dsz = int(w[2])
    assert np.array_equal(rt.buf[rt.total: rt.total + dsz], o.data)      # data area mirrored into the caller's buffer
    assert np.array_equal(rt.buf[: rt.total], o.code)                    # STORE_MEM words mirrored into the code area
    assert rt.reset(12345) == -1 and rt.reset(8000) == -2


def test_reload_params_live_patch(oracle_lib):
    """Host edits a gain PARAM word and a biquad bypass flag in the loaded program (dump-file workflow)."""
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 5, 400
    x = synth.pcm("noise", S, 2 * T, 2, 192000)
    ex = Executor(w, 192000, 2, S)
    y0 = ex.process(np.ascontiguousarray(x[:, :T]))
    w2 = w.copy()
    w2[57] = 1 << 27                      # LOAD_GAIN at word 54: gain literal at +3 -> 0.5
    ex.reload_params(w2)
    y1 = ex.process(np.ascontiguousarray(x[:, T:]))
    for s in range(S):
        o = oracle_lib.Oracle(w, 2, 192000, seed=0)
        a = o.process(x[s, :T])
        o.code[57] = 1 << 27
        b = o.process(x[s, T:])
        assert np.array_equal(y0[s], a) and np.array_equal(y1[s], b)


def _delay_param_program(with_biquad: bool):
    """LOAD_GAIN -> [BIQUADS] -> SAT0DB_TPDF -> DELAY(us from a PARAM word) -> STORE, twice, + TPDF_CALC.
    Returns (words, [word index of each delay PARAM])."""
    from oracle import wire
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.tpdf_calc(22)
    a.param()
    bq = a.biquad_sections([[wire.rbj_peak(48000, 900.0, 1.2, 1.4)], [wire.rbj_peak(48000, 3000.0, 0.8, 0.7)]]) if with_biquad else None
    d0 = a.delay_param(3000, 1500, 48000)
    d1 = a.delay_param(3000, 400, 48000)
    for ch, dp in ((0, d0), (1, d1)):
        a.load_gain(8 + ch, 0.6)
        if with_biquad:
            a.biquads(bq)
        a.sat0db_tpdf()
        a.delay(dp)
        a.store(ch)
    return a.end(), [d0, d1]


@pytest.mark.parametrize("with_biquad", [False, True])
@pytest.mark.parametrize("kernel", [KERNEL_GENERIC, KERNEL_AUTO])
def test_delay_time_patched_mid_stream(oracle_lib, with_biquad, kernel):
    """Host shortens / lengthens a delay PARAM mid-stream: the ring index can become stale (>= new length), which the
    reference uses once and then wraps (dsp_runtime.c:769-794).  All kernels must follow it bit for bit."""
    w, dps = _delay_param_program(with_biquad)
    fs, S, T = 48000, 5, 333
    x = synth.pcm("full", S, 3 * T, 2, fs)
    ex = Executor(w, fs, 2, S, seeds=np.arange(S, dtype=np.int32))
    ex.set_kernel(kernel)
    orcs = [oracle_lib.Oracle(w, 2, fs, seed=s) for s in range(S)]
    words = w.copy()
    for part, (us0, us1) in enumerate(((1500, 400), (200, 2900), (2500, 0))):
        for dp, us in zip(dps, (us0, us1)):
            words[dp] = (int(words[dp]) & ~0xFFFF) | us
        ex.reload_params(words)
        xs = np.ascontiguousarray(x[:, part * T:(part + 1) * T])
        y = ex.process(xs)
        for s in range(S):
            for dp, us in zip(dps, (us0, us1)):
                orcs[s].code[dp] = words[dp]
            assert np.array_equal(y[s], orcs[s].process(xs[s])), (part, s, ex.last_kernel)
    for s in (0, S - 1):
        st = ex.get_state(s)
        assert np.array_equal(st[: ex.data_size], orcs[s].data), (s, ex.last_kernel)


def test_state_roundtrip_and_reset():
    w = load_program("c5_mixer8x8_f2_192k")
    S, T = 4, 700
    x = synth.pcm("noise", S, T, 8, 192000)
    a = Executor(w, 192000, 2, S, seeds=[1, 2, 3, 4])
    ya = a.process(x)
    b = Executor(w, 192000, 2, S, seeds=[9, 9, 9, 9])
    b.process(np.ascontiguousarray(x[:, :123]))
    b.reset(seeds=[1, 2, 3, 4])
    assert np.array_equal(b.process(x), ya)
    # checkpoint/resume: copy stream 2's state into a fresh instance's stream 0
    c = Executor(w, 192000, 2, S, seeds=[1, 2, 3, 4])
    h = T // 2
    c.process(np.ascontiguousarray(x[:, :h]))
    d = Executor(w, 192000, 2, 1)
    d.set_state(0, c.get_state(2))
    yd = d.process(np.ascontiguousarray(x[2:3, h:]))
    assert np.array_equal(yd[0], ya[2, h:])


def test_edge_cases():
    w = load_program("c2_testrpi_xover_f2_192k")
    ex = Executor(w, 192000, 2, 1)
    assert ex.process(np.zeros((1, 0, 2), np.int32)).shape == (1, 0, 8)        # empty period
    y = ex.process(np.full((1, 64, 2), -(2 ** 31), np.int32))                  # most negative sample
    z = ex.process(np.full((1, 64, 2), 2 ** 31 - 1, np.int32))
    assert y.shape == z.shape == (1, 64, 8)
    one = Executor(w, 192000, 2, 1).process(synth.pcm("noise", 1, 1, 2, 192000))   # single frame
    assert one.shape == (1, 1, 8)


@pytest.mark.parametrize("prog,fs,S,T", [("c2_testrpi_xover_f2_192k", 192000, 4096, 4096),
                                          ("c5_mixer8x8_f2_192k", 192000, 4096, 2048)])
def test_full_width_batch_properties(oracle_lib, prog, fs, S, T):
    """BASELINE width (4096 streams): streams fed identical PCM and identical seeds must produce identical
    output; spot streams are checked bit-for-bit against the oracle; a checksum ties the rest together."""
    w = load_program(prog)
    ex = Executor(w, fs, 2, S)                                   # all seeds 0
    x1 = synth.pcm("noise", 1, T, ex.n_in, fs)
    x = np.ascontiguousarray(np.broadcast_to(x1, (S, T, ex.n_in)))
    y = ex.process(x)
    assert ex.last_kernel in ("chain", "mix")
    assert (y == y[0:1]).all()
    o = oracle_lib.Oracle(w, 2, fs, seed=0)
    assert np.array_equal(y[0], o.process(x1[0]))
    # distinct streams: spot-check a handful against the oracle
    ex2 = Executor(w, fs, 2, S, seeds=np.arange(S, dtype=np.int32))
    xs = synth.pcm("full", S, 512, ex.n_in, fs)
    ys = ex2.process(xs)
    for s in (0, 1, 31, 32, 1000, S - 1):
        o = oracle_lib.Oracle(w, 2, fs, seed=s)
        assert np.array_equal(ys[s], o.process(xs[s])), s


@pytest.mark.parametrize("stim", ["noise", "full"])
def test_chain3_full_width_distinct_streams(oracle_lib, stim):
    """The geometry the benchmark runs (4096 streams -> 28 per CTA, k_chain3) on 4096 DISTINCT streams: every lane class of a
    CTA (first, last live lane 27, first lane of the next CTA 28, middle, last stream) bit for bit against the oracle,
    outputs and state; 2048 frames so that AUTO takes k_chain3 (>= 1536)."""
    prog, fs, S, T = "c2_testrpi_xover_f2_192k", 192000, 4096, 2048
    w = load_program(prog)
    seeds = np.arange(S, dtype=np.int32) * 7 + 1
    ex = Executor(w, fs, 2, S, seeds=seeds)
    xs = synth.pcm(stim, S, T, ex.n_in, fs)
    ys = ex.process(xs)
    assert ex.last_kernel == "chain" and ex.last_chain_variant == 3
    for s in (0, 1, 5, 13, 26, 27, 28, 29, 55, 56, 2000, 2001, 4067, 4068, S - 1):
        o = oracle_lib.Oracle(w, 2, fs, seed=int(seeds[s]))
        assert np.array_equal(ys[s], o.process(xs[s])), s
        st = ex.get_state(s)
        assert np.array_equal(st[: ex.data_size], o.data), s
        assert np.array_equal(st[ex.aux_offset: ex.aux_offset + 7], o.aux()[:7]), s
    # no two distinct streams may produce the same output (a lane shadowing another would)
    sig = ys.reshape(S, -1)[:, 64 * 8: 64 * 8 + 64]
    assert len({row.tobytes() for row in sig}) == S


@pytest.mark.parametrize("prog,fmt", [("c4_fir4096_f3_48k", 3), ("c4s_fir_f3_multifs", 3), ("c4_fir4096_f2_48k", 2)])
def test_fir_kernel_periods_and_kernel_switches(oracle_lib, prog, fmt):
    """DSP_FIR through the time-parallel kernel: ALSA-period sized calls, calls shorter than the impulse, a switch to the
    generic interpreter and back (shared delay-line layout), all against one oracle run.  Float format 3 included:
    outputs AND the delay line must be identical (reference tap order, truncated products)."""
    from avdsp_b200 import KERNEL_FIR
    w = load_program(prog)
    fs, S, T = 48000, 5, 2600
    x = synth.pcm("full", S, T, 2, fs)
    seeds = np.arange(S, dtype=np.int32)
    ys, sts = oracle_run(oracle_lib, w, fmt, fs, x, seeds, 31)
    ex = Executor(w, fs, fmt, S, seeds=seeds)
    cuts = [0, 1, 3, 11, 300, 1324, 1325, 2349, T]
    parts = []
    for i, (c0, c1) in enumerate(zip(cuts, cuts[1:])):
        ex.set_kernel(KERNEL_GENERIC if i == 4 else KERNEL_FIR)
        parts.append(ex.process(np.ascontiguousarray(x[:, c0:c1])))
    y = np.concatenate(parts, axis=1)
    assert np.array_equal(y, ys), np.count_nonzero(y != ys)
    for s in (0, S - 1):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s
    # planar layout, one call
    ex2 = Executor(w, fs, fmt, S, seeds=seeds)
    yp = ex2.process(np.ascontiguousarray(x.transpose(0, 2, 1)), layout=PLANAR)
    assert ex2.last_kernel == "fir" and np.array_equal(yp.transpose(0, 2, 1), ys)


@pytest.mark.parametrize("prog,fmt", [("c4_fir4096_f2_48k", 2), ("c4_fir4096_f3_48k", 3)])
def test_fir_full_width_batch_properties(oracle_lib, prog, fmt):
    """C4 width (1024 streams, several time tiles per stream): spot streams bit-for-bit against the oracle, streams fed
    identical PCM give identical output, and a shifted copy of the input gives the shifted output (time invariance
    across tile boundaries; the FIR has no other state)."""
    w = load_program(prog)
    fs, S, T = 48000, 1024, 6144
    xs = synth.pcm("noise", S, T, 2, fs)
    xs[1] = xs[0]
    D = 1000
    xs[2, :D] = 0; xs[2, D:] = xs[0, : T - D]
    ex = Executor(w, fs, fmt, S)
    y = ex.process(xs)
    assert ex.last_kernel == ("fir_tc" if fmt == 2 else "fir")      # fixed point: bit-exact tensor-core limb GEMM
    assert np.array_equal(y[1], y[0])
    assert np.array_equal(y[2, D:], y[0, : T - D]) and not y[2, :D].any()
    for s in (0, 3, 517, S - 1):
        o = oracle_lib.Oracle(w, fmt, fs, seed=0)
        assert np.array_equal(y[s], o.process(xs[s])), s
        assert np.array_equal(ex.get_state(s)[: ex.data_size], o.data), s


def test_fir_tensor_core_int8_limbs_bit_exact(oracle_lib):
    """DSP_FORMAT 2 FIR as a Toeplitz GEMM on tcgen05 (kind::i8, four 8-bit limbs per operand, int32 accumulators in
    TMEM, recombined in wrapping int64): must be BIT-EXACT, outputs and delay line, for ragged stream counts (partial
    64-stream tile), calls shorter than one 128-output block, full-scale input that saturates SAT0DB, and when
    alternated with the scalar-pipe kernel and the interpreter."""
    from avdsp_b200 import KERNEL_FIR, KERNEL_FIR_TC
    for prog, S, T in (("c4_fir4096_f2_48k", 70, 1500), ("c4s_fir_f2_multifs", 5, 700)):
        w = load_program(prog)
        fs = 48000
        x = synth.pcm("full", S, T, 2, fs)
        seeds = np.arange(S, dtype=np.int32)
        ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 24)
        ex = Executor(w, fs, 2, S, seeds=seeds, dither=24)
        ex.set_kernel(KERNEL_FIR_TC)
        y = ex.process(x)
        assert ex.last_kernel == "fir_tc"
        assert np.array_equal(y, ys), (prog, np.count_nonzero(y != ys))
        for s in (0, S - 1):
            assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), (prog, s)
        ex2 = Executor(w, fs, 2, S, seeds=seeds, dither=24)
        cuts = [0, 1, 130, 131, 700, T] if T > 700 else [0, 1, 130, 131, T]
        kern = [KERNEL_FIR_TC, KERNEL_FIR, KERNEL_FIR_TC, KERNEL_GENERIC, KERNEL_FIR_TC]
        parts = []
        for i, (c0, c1) in enumerate(zip(cuts, cuts[1:])):
            ex2.set_kernel(kern[i])
            parts.append(ex2.process(np.ascontiguousarray(x[:, c0:c1])))
        assert np.array_equal(np.concatenate(parts, axis=1), ys), prog
        assert np.array_equal(ex2.get_state(S - 1), expected_state(ex2, sts[S - 1])), prog


def test_fir_tensor_core_tf32_stated_tolerance(oracle_lib):
    """DSP_FORMAT 3 FIR as a 3xTF32 Toeplitz GEMM (opt-in): NOT the reference's summation order.  Stated tolerance:
    |y - reference| <= 2^-16 of full scale (-96 dBFS) on s.31 outputs; the delay line (exact input samples) must be
    identical.  The exact-order float kernel stays the AUTO choice."""
    from avdsp_b200 import KERNEL_FIR_TC
    w = load_program("c4_fir4096_f3_48k")
    fs, S, T = 48000, 40, 1100
    x = synth.pcm("noise", S, T, 2, fs)
    ys, sts = oracle_run(oracle_lib, w, 3, fs, x, np.zeros(S, np.int32), 31)
    ex = Executor(w, fs, 3, S)
    ex.set_kernel(KERNEL_FIR_TC)
    y = np.concatenate([ex.process(np.ascontiguousarray(x[:, :333])), ex.process(np.ascontiguousarray(x[:, 333:]))], axis=1)
    assert ex.last_kernel == "fir_tc"
    err = np.abs(y.astype(np.int64) - ys.astype(np.int64)).max()
    print("3xTF32 FIR: max |err| =", err, "LSB of s.31 =", err / 2.0 ** 31, "FS")
    assert err <= 2 ** 15, err
    assert np.abs(ys).max() > 2 ** 28                      # the comparison is not vacuous
    for s in (0, S - 1):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s


def _uniform_delay_program():
    """Four outputs, every path LOAD_MUX/LOAD_GAIN -> SAT0DB_TPDF_GAIN -> DELAY(us from a PARAM word) -> STORE + TPDF_CALC:
    a 'uniform' program, i.e. what the per-stream mix kernel takes.  Returns (words, delay PARAM word indices)."""
    from oracle import wire
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.tpdf_calc(23)
    a.param()
    dps = [a.delay_param(4000, us, 48000) for us in (1500, 400, 0, 3900)]
    tabs = [a.mux_table([(8 + (k + j) % 4, 0.3 - 0.1 * j) for j in range(3)]) for k in range(4)]
    for ch in range(4):
        a.load_mux(tabs[ch])
        a.sat0db_tpdf_gain(0.9 - 0.1 * ch)
        a.delay(dps[ch])
        a.store(ch)
    return a.end(), dps


def test_stream_mix_kernel_stale_delays_and_state(oracle_lib):
    """Per-stream mix kernel (uniform program): delay times patched mid-stream (stale ring indices, zero-length delays),
    many streams (one CTA per stream: in-kernel dither generation + in-kernel state write-back) and few streams
    (time segments + the separate tail kernel), against the oracle: outputs and every state word."""
    w, dps = _uniform_delay_program()
    fs = 48000
    for S, T in ((900, 2304), (3, 9000), (5, 333)):
        x = synth.pcm("full", S, 3 * T, 4, fs)
        seeds = np.arange(S, dtype=np.int32)
        ex = Executor(w, fs, 2, S, seeds=seeds)
        pick = sorted({0, 1, S // 2, S - 1})
        orcs = {s: oracle_lib.Oracle(w, 2, fs, seed=s) for s in pick}
        words = w.copy()
        for part, uss in enumerate(((1500, 400, 0, 3900), (200, 3000, 700, 0), (3999, 1, 5, 100))):
            for dp, us in zip(dps, uss):
                words[dp] = (int(words[dp]) & ~0xFFFF) | us
            ex.reload_params(words)
            xs = np.ascontiguousarray(x[:, part * T:(part + 1) * T])
            y = ex.process(xs)
            assert ex.last_kernel == "mix"
            for s in pick:
                for dp, us in zip(dps, uss):
                    orcs[s].code[dp] = words[dp]
                assert np.array_equal(y[s], orcs[s].process(xs[s])), (S, part, s)
        for s in pick:
            st = ex.get_state(s)
            assert np.array_equal(st[: ex.data_size], orcs[s].data), (S, s)
            assert np.array_equal(st[ex.aux_offset: ex.aux_offset + 7], orcs[s].aux()[:7]), (S, s)


def test_c1_ten_seconds_of_stereo_pcm(oracle_lib):
    """BASELINE.json configs[0]: the stereo 2-way LR4 crossover (crossover2x2lfe, 3 io hand-offs through MEM words) over
    10 s of synthetic 48 kHz stereo PCM, one stream: bit-exact plumbing against the oracle, outputs and data area."""
    w = load_program("c1_crossover2x2lfe_f2_48k")
    fs, T = 48000, 480000
    x = synth.pcm("noise", 1, T, 2, fs)
    o = oracle_lib.Oracle(w, 2, fs, seed=0, dither=31)
    yo = o.process(x[0])
    ex = Executor(w, fs, 2, 1, seeds=[0], dither=31)
    y = np.concatenate([ex.process(np.ascontiguousarray(x[:, c0:c0 + 96000])) for c0 in range(0, T, 96000)], axis=1)
    assert np.array_equal(y[0], yo)
    assert np.array_equal(ex.get_state(0)[: ex.data_size], o.data)


def _mem_handoff_program():
    """What oktodac-style crossovers do: core 1 copies inputs through (LOAD_STORE, one slot twice), computes the dither and
    runs shared pre-filters into MEM words; later cores continue those cascades per output.  One output is stored twice
    (the later path wins), one MEM word feeds two consumers."""
    from oracle import wire
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core()
    a.load_store([(8, 4), (9, 5), (8, 5)])
    a.tpdf_calc(24)
    a.param()
    m0 = a.mem_location(); m1 = a.mem_location()
    pre0 = a.biquad_sections([[wire.rbj_peak(48000, 300.0, 0.9, 1.3)], [wire.rbj_peak(48000, 5000.0, 0.7, 0.8)]])
    pre1 = a.biquad_sections([[wire.rbj_peak(48000, 120.0, 1.1, 1.2)]] * 2)
    lo = a.biquad_sections([[wire.rbj_peak(48000, 800.0, 0.7, 0.5)]] * 2)
    hi = a.biquad_sections([[wire.rbj_peak(48000, 2500.0, 0.7, 1.6)]] * 4)
    d0 = a.delay_param(2000, 700, 48000)
    a.load_gain(8, 0.45); a.biquads(pre0); a.store_mem(m0)
    a.load_gain(9, 0.45); a.biquads(pre1); a.store_mem(m1)
    a.core()
    a.load_mem(m0); a.biquads(lo); a.sat0db_tpdf(); a.delay(d0); a.store(0); a.store(2)
    a.load_mem(m0); a.biquads(hi); a.sat0db(); a.store(1)
    a.core()
    a.load_mem(m1); a.sat0db_tpdf_gain(0.9); a.store(3)
    a.load_mem(m1); a.biquads(lo); a.sat0db(); a.store(2)        # overwrites core 2's STORE 2
    return a.end()


def test_chain_kernel_inlines_mem_handoffs(oracle_lib):
    """LOAD_STORE pass-through paths, TPDF_CALC behind them, cascades handed between cores through MEM words, dead stores:
    the chain kernel must take the program and match the oracle bit for bit -- outputs, biquad/delay state, and the MEM
    words the producers leave in the (mirrored) code area."""
    w = _mem_handoff_program()
    fs, S, T = 48000, 37, 700
    seeds = np.arange(S, dtype=np.int32) + 3
    x = synth.pcm("full", S, T, 2, fs)
    ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 31)
    ex = Executor(w, fs, 2, S, seeds=seeds)
    cuts = [0, 1, 33, 400, T]
    y = np.concatenate([ex.process(np.ascontiguousarray(x[:, c0:c1])) for c0, c1 in zip(cuts, cuts[1:])], axis=1)
    assert ex.last_kernel == "chain", ex.trace
    assert np.array_equal(y, ys), np.count_nonzero(y != ys)
    for s in (0, 5, S - 1):
        got, exp = ex.get_state(s), expected_state(ex, sts[s])
        assert np.array_equal(got, exp), (s, np.nonzero(got != exp)[0][:10])
    g = Executor(w, fs, 2, S, seeds=seeds); g.set_kernel(KERNEL_GENERIC)
    assert np.array_equal(g.process(x), ys)


def test_delay_first_paths_with_patched_delays(oracle_lib):
    """cascade -> DELAY -> SAT0DB[_TPDF][_GAIN] -> STORE (the osx/dacdiy1.bin shape): the ring holds the low word of the Q59
    accumulator.  Delay times are patched mid-stream (stale ring indices, a delay dropping to zero); chain kernel and
    interpreter against the oracle, outputs and state."""
    from oracle import wire
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.tpdf_calc(22)
    a.param()
    bq = a.biquad_sections([[wire.rbj_peak(48000, 900.0, 1.2, 1.4)], [wire.rbj_peak(48000, 3000.0, 0.8, 0.7)]])
    dps = [a.delay_param(3000, us, 48000) for us in (1500, 400, 2900, 50)]
    for ch in range(4):
        a.load_gain(8 + (ch & 1), 0.6 - 0.1 * ch)
        a.biquads(bq)
        a.delay(dps[ch])
        [lambda: a.sat0db_tpdf_gain(0.9), a.sat0db_tpdf, lambda: a.sat0db_gain(0.8), a.sat0db][ch]()
        a.store(ch)
    w = a.end()
    fs, S, T = 48000, 6, 333
    x = synth.pcm("full", S, 3 * T, 2, fs)
    for kernel in (KERNEL_AUTO, KERNEL_GENERIC):
        ex = Executor(w, fs, 2, S, seeds=np.arange(S, dtype=np.int32))
        ex.set_kernel(kernel)
        orcs = [oracle_lib.Oracle(w, 2, fs, seed=s) for s in range(S)]
        words = w.copy()
        for part, uss in enumerate(((1500, 400, 2900, 50), (200, 2900, 0, 700), (2500, 0, 1000, 10))):
            for dp, us in zip(dps, uss):
                words[dp] = (int(words[dp]) & ~0xFFFF) | us
            ex.reload_params(words)
            xs = np.ascontiguousarray(x[:, part * T:(part + 1) * T])
            y = ex.process(xs)
            if kernel == KERNEL_AUTO:
                assert ex.last_kernel == "chain", ex.trace
            for s in range(S):
                for dp, us in zip(dps, uss):
                    orcs[s].code[dp] = words[dp]
                assert np.array_equal(y[s], orcs[s].process(xs[s])), (part, s, ex.last_kernel)
        for s in (0, S - 1):
            assert np.array_equal(ex.get_state(s)[: ex.data_size], orcs[s].data), (s, ex.last_kernel)


@pytest.mark.parametrize("prog,fs", [("c5_mixer8x8_f2_192k", 192000), ("c2_testrpi_xover_f2_192k", 192000), ("c4_fir4096_f2_48k", 48000)])
def test_ranges_on_two_cuda_streams_equal_one_call(prog, fs):
    """avdsp_b200_process_range with firstStream != 0 on two CUDA streams at once: launches of one instance share scratch
    (dither rows, FIR workspace, PRNG jump matrices), the library orders them itself (include/avdsp_b200.h)."""
    import torch
    w = load_program(prog)
    S, T = 96, 1024
    seeds = np.arange(S, dtype=np.int32) + 3
    a = Executor(w, fs, 2, S, seeds=seeds)
    b = Executor(w, fs, 2, S, seeds=seeds)
    x = torch.from_numpy(synth.pcm("full", S, T, a.n_in, fs)).cuda()
    ref = a.process(x)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    cut = 40
    for rep in range(2):                        # second round: continues the state, other split
        lo, hi = (cut, S - cut) if rep == 0 else (S - cut, cut)
        if rep == 1:
            ref = a.process(x)
            torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            y1 = b.process_range(x[:lo].contiguous(), 0, stream=s1)
        with torch.cuda.stream(s2):
            y2 = b.process_range(x[lo:].contiguous(), lo, stream=s2)
        torch.cuda.synchronize()
        assert torch.equal(torch.cat([y1, y2]), ref), rep
    for s in (0, cut - 1, cut, S - 1):
        assert np.array_equal(a.get_state(s), b.get_state(s))


@pytest.mark.parametrize("replicate", [1, 3])
def test_multi_device_instance_fans_host_buffers_out(oracle_lib, monkeypatch, replicate):
    """avdsp_b200_create_multi (SURVEY.md 8b deviceMask): one C call over every GPU of the box.  On a one-GPU box the shards
    are replicated on device 0 (AVDSP_B200_MULTI_REPLICATE) so that the range arithmetic is exercised all the same."""
    import torch
    ndev = torch.cuda.device_count()
    monkeypatch.setenv("AVDSP_B200_MULTI_REPLICATE", str(replicate))
    w = load_program("c2_testrpi_xover_f2_192k")
    fs, S, T = 192000, 37, 700
    seeds = np.arange(S, dtype=np.int32) * 5 + 2
    ex = Executor(w, fs, 2, S, seeds=seeds, dither=24, devices=list(range(ndev)))
    sh = ex.shards()
    assert len(sh) == ndev * replicate and sh[0][1] == 0 and sum(n for _, _, n, _ in sh) == S
    assert all(a[1] + a[2] == b[1] for a, b in zip(sh, sh[1:]))                  # contiguous ranges
    assert max(n for _, _, n, _ in sh) - min(n for _, _, n, _ in sh) <= 1         # balanced
    x = ex.alloc_pcm(T, ex.n_in)
    y = ex.alloc_pcm(T, ex.n_out)
    x[:] = synth.pcm("full", S, T, ex.n_in, fs)
    ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 24)
    ex.process(x[:, :300].copy(), out=None)                                         # plain (unplaced) memory works too
    ex.reset(seeds=seeds, dither=24)
    ex.process(x, out=y)
    assert np.array_equal(y, ys)
    for s in (0, sh[-1][1], S - 1):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s
    z = np.zeros_like(y)
    ex.copy_only(x, z)                                                              # same DMA schedule, no kernel: state untouched
    assert np.array_equal(ex.get_state(S - 1), expected_state(ex, sts[S - 1]))
    with pytest.raises(AvdspError):
        ex.process(torch.zeros((S, 8, ex.n_in), dtype=torch.int32, device="cuda"))  # device buffers belong to one GPU
    # S16 periods through the multi instance
    raw = (x[:, :64] >> 16).astype(np.int16)
    ex.reset(seeds=seeds, dither=24)
    one = Executor(w, fs, 2, S, seeds=seeds, dither=24)
    assert np.array_equal(ex.process_pcm(raw, 1, 64), one.process_pcm(raw, 1, 64))
