"""k_chain3 (kernel_chain3.cu: a biquad cascade, or a part of one, per lane, lane = stream) against the CPU oracle through the
C ABI.  Fixed point and the float class (DSP_FORMAT 3 / 5), bar = BIT-EXACT: outputs and every state word (biquad accumulators
and histories, delay rings + indices, PRNG / TPDF words).  The kernel is forced with KERNEL_CHAIN_V3 and given several streams per CTA
through AVDSP_B200_NS3 (AUTO picks it by itself only at batch width, which test_gpu_parity's 4096-stream tests cover).
"""
import numpy as np
import pytest

from conftest import load_program
from avdsp_b200 import Executor, AvdspError, synth, KERNEL_GENERIC, KERNEL_CHAIN_V2, KERNEL_CHAIN_V3, KERNEL_AUTO

from test_gpu_parity import expected_state, oracle_run, _delay_param_program

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["4", "8", "2"], ids=["parts4", "whole", "parts2"])
def _several_streams_per_cta(monkeypatch, request):
    """Both shapes of the kernel: cascades cut into parts of <= 4 (or 2) sections that run as separate warps a tile behind each
    other, and whole cascades per lane.  (Where the parts do not fit 16 warps the geometry falls back to whole cascades.)"""
    monkeypatch.setenv("AVDSP_B200_NS3", "5")          # read when an Executor plans its geometry
    monkeypatch.setenv("AVDSP_B200_PART3", request.param)


def _v3(w, fs, S, **kw):
    ex = Executor(w, fs, 2, S, **kw)
    ex.set_kernel(KERNEL_CHAIN_V3)
    return ex


@pytest.mark.parametrize("stim", ["full", "noise", "impulse"])
def test_c3_sixteen_section_cascades_in_four_parts(oracle_lib, stim):
    """C3 (16-section PEQ per channel, plain SAT0DB): the cascade runs as four part warps, each a tile behind the previous one
    (lag 108 frames); where the test fixture asks for whole cascades or parts of two the geometry does not fit and v3 declines."""
    w, fs = load_program("c3_peq16_f2_48k"), 48000
    S, T = 23, 777
    seeds = np.arange(S, dtype=np.int32)
    ex = Executor(w, fs, 2, S, seeds=seeds, dither=31)
    ex.set_kernel(KERNEL_CHAIN_V3)
    x = synth.pcm(stim, S, T, ex.n_in, fs)
    try:
        a = ex.process(x[:, :300])
    except AvdspError:
        assert "chain kernel v3 geometry" not in ex.trace          # parts of 2 (8 per chain) or whole 16-section cascades: no v3 shape
        return
    b = ex.process(x[:, 300:])                                     # a second call: fill / drain of the four-deep part pipeline
    assert ex.last_chain_variant == 3
    ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 31)
    assert np.array_equal(np.concatenate([a, b], axis=1), ys)
    for s in (0, 4, 5, S - 1):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s


def _float_state_close(got, exp):
    """float class of the chain kernels: products are mul.rz.ftz.f32, which is dspMulFloatFloat (dsp_ieee754.h:336-375) except next to
    the underflow threshold and beyond the binary32 range; streams that come near either are flagged by the cascades and re-executed
    by the interpreter from a state snapshot (avdsp_dev.cuh fltGuard, api.cu launchRun).  So: bit for bit, state included."""
    return np.array_equal(got, exp)


@pytest.mark.parametrize("kernel", [KERNEL_CHAIN_V2, KERNEL_CHAIN_V3, KERNEL_AUTO])
def test_float_class_decay_into_the_underflow_range(oracle_lib, kernel):
    """An impulse, then silence: the cascades' state decays through 2^-64 ... 2^-126 to zero.  There the reference flushes products
    on the exponent sum before normalisation and its float -> s.31 conversion shifts by (127 - exponent) mod 32, i.e. it emits
    pseudo-random samples for tiny values: every bit of that is reproduced (second pass of the interpreter on the flagged
    streams), outputs and state -- the accumulators end up parked next to 2^-126 for good; streams 3.. keep playing noise and
    stay on the fast path."""
    from oracle import wire
    fs = 48000
    a = wire.Asm(fmt=3, fmin=fs, fmax=fs)
    a.core(); a.tpdf_calc(24); a.param()
    e1 = a.biquad_sections([[wire.rbj_peak(fs, 6000.0 + 1500 * k, 0.5, 1.3)] for k in range(8)])      # low Q: the state decays a decade in a few frames
    e2 = a.biquad_sections([[wire.rbj_peak(fs, 9000.0 + 1000 * k, 0.6, 0.8)] for k in range(4)])
    a.load_gain(8, 0.8); a.biquads(e1); a.sat0db(); a.store(0)
    a.load(9); a.biquads(e2); a.sat0db_tpdf(); a.store(1)
    w = a.end()
    S, T = 9, 3300
    seeds = np.arange(S, dtype=np.int32)
    ex = Executor(w, fs, 3, S, seeds=seeds, dither=24)
    ex.set_kernel(kernel)
    x = synth.pcm("noise", S, T, ex.n_in, fs)
    x[:3, 40:] = 0                                               # three streams fall silent after 40 frames
    x[1, 2500:] = synth.pcm("noise", 1, T - 2500, ex.n_in, fs)[0]  # ... one of them starts again
    try:
        y = np.concatenate([ex.process(x[:, :1700]), ex.process(x[:, 1700:])], axis=1)
    except AvdspError:
        assert kernel == KERNEL_CHAIN_V3 and "chain kernel v3 geometry" not in ex.trace   # the fixture's other part sizes
        return
    ys, sts = oracle_run(oracle_lib, w, 3, fs, x, seeds, 24)
    ex_field = (sts[0][0][: ex.data_size].view(np.uint32) >> 23) & 255
    assert np.any((ex_field >= 1) & (ex_field < 63)), "the silent stream's cascade state has not reached the underflow range"
    assert np.array_equal(y, ys), (np.count_nonzero(y != ys), sorted(set(np.nonzero(y != ys)[0])))
    for s in range(S):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s


@pytest.mark.parametrize("stim", ["noise", "impulse", "sine", "full"])
def test_float_class_dsp_format_3(oracle_lib, stim):
    """DSP_FORMAT 3 (float ALU, int32 samples) on k_chain3: C3's 16-section cascades as four part warps, and a shorter program
    with TPDF finish, LOAD without gain and a delay line; s.31 outputs identical to the reference's float arithmetic.
    "full" drives the cascades past 0 dB: interior tiles park unsaturated floats in the post ring, the store warps and the
    delay line's write-back saturate (dspSaturateFloat0db is idempotent)."""
    from oracle import wire
    fs = 48000
    a = wire.Asm(fmt=3, fmin=fs, fmax=fs)
    a.core(); a.tpdf_calc(24); a.param()
    e1 = a.biquad_sections([[wire.rbj_peak(fs, 300 * (k + 1), 1.0 + 0.3 * k, 1.4 if k & 1 else 0.7)] for k in range(5)])
    e2 = a.biquad_sections([[wire.rbj_peak(fs, 900 * (k + 1), 2.0, 0.8)] for k in range(3)])
    dl = a.delay_param(2000, 700, fs)
    a.load_gain(8, 0.8); a.biquads(e1); a.sat0db_tpdf(); a.delay(dl); a.store(0)
    a.load(9); a.biquads(e2); a.sat0db(); a.store(1)
    own = a.end()
    for w, S, T in ((load_program("c3_peq16_f3_48k"), 23, 777), (own, 11, 500)):
        seeds = np.arange(S, dtype=np.int32) + 5
        ex = Executor(w, fs, 3, S, seeds=seeds, dither=24)
        ex.set_kernel(KERNEL_CHAIN_V3)
        x = synth.pcm(stim, S, T, ex.n_in, fs)
        try:
            y1 = ex.process(x[:, :260])
        except AvdspError:
            assert "chain kernel v3 geometry" not in ex.trace      # fixture shapes the float form does not have (parts > 4 sections)
            continue
        y2 = ex.process(x[:, 260:])
        assert ex.last_chain_variant == 3
        ys, sts = oracle_run(oracle_lib, w, 3, fs, x, seeds, 24)
        assert np.array_equal(np.concatenate([y1, y2], axis=1), ys)
        for s in (0, 4, 5, S - 1):
            got, exp = ex.get_state(s), expected_state(ex, sts[s])
            dd = np.nonzero(got != exp)[0]
            assert _float_state_close(got, exp), (s, dd[:10], got[dd[:10]].view(np.float32), exp[dd[:10]].view(np.float32), got[dd[:10]], exp[dd[:10]])


@pytest.mark.parametrize("prog,fs", [("c2_testrpi_xover_f2_192k", 192000)])
@pytest.mark.parametrize("stim", ["full", "noise", "impulse", "sine"])
def test_c2_ragged_batch_bit_exact(oracle_lib, prog, fs, stim):
    """C2 program (8 cascades of 4/4/6/8/6/6/8/6 sections, TPDF dither on two of them, six delays), ragged sizes: 37 streams
    at 5 per CTA (last CTA: 2), 333 frames (partial last tile).  Full-scale noise saturates sections (replayed tiles)."""
    w = load_program(prog)
    S, T = 37, 333
    seeds = np.arange(S, dtype=np.int32) * 7 + 1
    ex = _v3(w, fs, S, seeds=seeds, dither=24)
    x = synth.pcm(stim, S, T, ex.n_in, fs)
    ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 24)
    y = ex.process(x)
    assert ex.last_kernel == "chain" and ex.last_chain_variant == 3, ex.trace
    assert np.count_nonzero(y != ys) == 0, f"{np.count_nonzero(y != ys)}/{y.size} samples differ"
    for s in (0, 1, 4, 5, 35, S - 1):
        got, exp = ex.get_state(s), expected_state(ex, sts[s])
        diff = np.nonzero(got != exp)[0]
        assert diff.size == 0, f"stream {s}: state words {diff[:8]} differ"


def test_any_split_into_periods_and_kernel_switches(oracle_lib):
    """ALSA periods of any size (1 frame, odd sizes, around tile boundaries) and alternating with the generic and the v2
    chain kernels mid-stream: identical output and state (one state layout for all kernels)."""
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 11, 1500
    x = synth.pcm("full", S, T, 2, 192000)
    seeds = np.arange(S, dtype=np.int32)
    a = Executor(w, 192000, 2, S, seeds=seeds); a.set_kernel(KERNEL_GENERIC)
    ya = a.process(x)
    b = _v3(w, 192000, S, seeds=seeds)
    cuts = [0, 1, 2, 33, 65, 66, 130, 577, 1024, 1056, T]
    yb = np.concatenate([b.process(np.ascontiguousarray(x[:, c0:c1])) for c0, c1 in zip(cuts, cuts[1:])], axis=1)
    assert b.last_chain_variant == 3
    assert np.array_equal(ya, yb)
    c = Executor(w, 192000, 2, S, seeds=seeds)
    parts = []
    for i, (c0, c1) in enumerate(zip(cuts, cuts[1:])):
        c.set_kernel((KERNEL_CHAIN_V3, KERNEL_GENERIC, KERNEL_CHAIN_V2)[i % 3])
        parts.append(c.process(np.ascontiguousarray(x[:, c0:c1])))
    assert np.array_equal(ya, np.concatenate(parts, axis=1))
    for s in (0, 4, 5, S - 1):
        assert np.array_equal(a.get_state(s), b.get_state(s)), s
        assert np.array_equal(a.get_state(s), c.get_state(s)), s


def _odd_program():
    """Cascades of 1, 3, 5 and 7 sections; sources LOAD (no gain: the cascade sees X >> 28), LOAD_GAIN 1.0, 0.35 and -1.7;
    SAT0DB and SAT0DB_TPDF finishes; fixed delays of 0, 7 and 40 samples; paths with two and four STOREs; output 0 is stored by
    the first path and again by the last one (the later STORE wins, the first path still runs for its state)."""
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
    return a.end(), fs


@pytest.mark.parametrize("stim", ["full", "noise"])
def test_odd_cascade_lengths_sources_and_finishes(oracle_lib, stim):
    w, fs = _odd_program()
    S, T = 13, 777
    seeds = np.arange(S, dtype=np.int32) + 3
    ex = _v3(w, fs, S, seeds=seeds, dither=24)
    x = synth.pcm(stim, S, T, ex.n_in, fs)
    ys, sts = oracle_run(oracle_lib, w, 2, fs, x, seeds, 24)
    y = ex.process(x)
    assert ex.last_chain_variant == 3, ex.trace
    assert np.array_equal(y, ys), f"{np.count_nonzero(y != ys)}/{y.size} samples differ"
    for s in (0, 4, 5, S - 1):
        assert np.array_equal(ex.get_state(s), expected_state(ex, sts[s])), s


def test_delay_times_patched_mid_stream(oracle_lib):
    """Delay PARAMs shortened / lengthened / zeroed between calls: stale ring indices (used once, then wrapped,
    dsp_runtime.c:769-794) -- frame 0 of such a call takes the exact path and swaps with ring[idx0] in the state block."""
    w, dps = _delay_param_program(True)
    fs, S, T = 48000, 7, 333
    x = synth.pcm("full", S, 3 * T, 2, fs)
    ex = _v3(w, fs, S, seeds=np.arange(S, dtype=np.int32))
    orcs = [oracle_lib.Oracle(w, 2, fs, seed=s) for s in range(S)]
    words = w.copy()
    for part, (us0, us1) in enumerate(((1500, 400), (200, 2900), (2500, 0))):
        for dp, us in zip(dps, (us0, us1)):
            words[dp] = (int(words[dp]) & ~0xFFFF) | us
        ex.reload_params(words)
        xs = np.ascontiguousarray(x[:, part * T:(part + 1) * T])
        y = ex.process(xs)
        assert ex.last_chain_variant == 3, ex.trace
        for s in range(S):
            for dp, us in zip(dps, (us0, us1)):
                orcs[s].code[dp] = words[dp]
            assert np.array_equal(y[s], orcs[s].process(xs[s])), (part, s, ex.last_kernel, ex.last_chain_variant)
    for s in (0, S - 1):
        assert np.array_equal(ex.get_state(s)[: ex.data_size], orcs[s].data), s


def test_foreign_state_block_with_incoherent_histories():
    """The fast path keeps only the y histories inside a cascade (x history of section k+1 == y history of section k for every
    state the reference can reach).  A state block that violates this (set_state with arbitrary words) must still be honoured:
    same result as the generic interpreter started from the same block."""
    w = load_program("c2_testrpi_xover_f2_192k")
    S, T = 6, 200
    rng = np.random.default_rng(5)
    x = synth.pcm("noise", S, T, 2, 192000)
    a = Executor(w, 192000, 2, S); a.set_kernel(KERNEL_GENERIC)
    b = _v3(w, 192000, S)
    blk = a.get_state(0)
    for s in range(S):
        st = blk.copy()
        st[2:50] = rng.integers(-2 ** 20, 2 ** 20, 48)          # biquad words of the two core-1 cascades: incoherent on purpose
        a.set_state(s, st); b.set_state(s, st)
    ya, yb = a.process(x), b.process(x)
    assert b.last_chain_variant == 3
    assert np.array_equal(ya, yb)
    for s in range(S):
        assert np.array_equal(a.get_state(s), b.get_state(s)), s


def test_v3_refuses_what_it_cannot_run():
    """Program shapes outside v3 (mixers, gains behind the cascade, the double-ALU format 4) are refused when v3 is forced and
    run on the other kernels under AUTO; short calls stay on v2 under AUTO even where v3 could run them."""
    for prog, fmt, fs, kern in (("c5_mixer8x8_f2_192k", 2, 192000, "mix"), ("c1_crossover2x2lfe_f2_48k", 2, 48000, "dag"), ("c3_peq16_f4_48k", 4, 48000, "generic")):
        ex = Executor(load_program(prog), fs, fmt, 8)
        x = (synth.pcm_float if fmt >= 5 else synth.pcm)("noise", 8, 64, ex.n_in, fs)
        ex.set_kernel(KERNEL_CHAIN_V3)
        with pytest.raises(AvdspError):
            ex.process(x)
        ex.set_kernel(KERNEL_AUTO)
        ex.process(x)
        assert ex.last_kernel == kern
    for prog, fmt, fs in (("c3_peq16_f2_48k", 2, 48000), ("c3_peq16_f3_48k", 3, 48000)):
        ex = Executor(load_program(prog), fs, fmt, 8)
        ex.process(synth.pcm("noise", 8, 64, ex.n_in, fs))
        assert ex.last_kernel == "chain" and ex.last_chain_variant == 2


@pytest.mark.parametrize("T", [257, 96, 1000])
def test_planar_layout_device_buffers_and_bulk_copies(oracle_lib, T):
    """Every way the PCM can reach the kernel: planar [stream][channel][frame] host buffers (input tiles staged with plain
    copies, lane = frame stores), interleaved device buffers whose rows are 16-byte friendly (bulk copies of full tiles, plain
    copy of a partial last tile) and rows that are not (T odd: plain copies only)."""
    import torch
    from avdsp_b200 import PLANAR
    w = load_program("c2_testrpi_xover_f2_192k")
    S = 33
    seeds = np.arange(S, dtype=np.int32) + 11
    x = synth.pcm("full", S, T, 2, 192000)
    ys, sts = oracle_run(oracle_lib, w, 2, 192000, x, seeds, 31)
    ex = _v3(w, 192000, S, seeds=seeds)
    yp = ex.process(np.ascontiguousarray(x.transpose(0, 2, 1)), layout=PLANAR)
    assert ex.last_chain_variant == 3
    assert np.array_equal(yp.transpose(0, 2, 1), ys)
    assert np.array_equal(ex.get_state(S - 1), expected_state(ex, sts[S - 1]))
    ex2 = _v3(w, 192000, S, seeds=seeds)
    yd = ex2.process(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert ex2.last_chain_variant == 3
    assert np.array_equal(yd.cpu().numpy(), ys)
    ex3 = _v3(w, 192000, S, seeds=seeds)
    yq = ex3.process(torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).cuda(), layout=PLANAR)
    torch.cuda.synchronize()
    assert np.array_equal(yq.cpu().numpy().transpose(0, 2, 1), ys)
    assert np.array_equal(ex3.get_state(0), expected_state(ex3, sts[0]))
