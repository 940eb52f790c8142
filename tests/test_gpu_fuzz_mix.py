"""Randomised programs WITHOUT biquads (mixers, routers, delay / dither programs) and FIR programs at random batch shapes through
AUTO's choice -- k_mix_stream / k_mix_main (+ the segmented PRNG when few streams carry many frames), k_fir, k_firtc -- against
the oracle, bit for bit in outputs and every state word (fixed point; the float formats ride along on whatever AUTO picks)."""
import numpy as np
import pytest

from avdsp_b200 import Executor, synth, INTERLEAVED, PLANAR
from oracle import wire
from test_gpu_parity import expected_state

pytestmark = pytest.mark.gpu


def random_mix_program(rng, fs=48000, fmt=2):
    a = wire.Asm(fmt=fmt, fmin=fs, fmax=fs)
    n_in = int(rng.choice([1, 2, 4, 8]))
    n_out = int(rng.choice([1, 2, 3, 4, 8]))
    ins = [8 + k for k in range(n_in)]
    uniform = rng.random() < 0.4                          # every output a LOAD_MUX over all inputs with the same finish: the dense-matrix form
    fin = int(rng.integers(0, 5))
    # outputs -> cores (tables live in the core's PARAM section, in front of its code)
    cores, cur = [], []
    for o in range(n_out):
        if cur and rng.random() < 0.25:
            cores.append(cur); cur = []
        cur.append(o)
    cores.append(cur)
    for ci, outs in enumerate(cores):
        a.core()
        if ci == 0 and rng.random() < 0.8:
            a.tpdf_calc(int(rng.choice([16, 20, 24, 31])))
        a.param()
        tabs = {}
        for o in outs:
            sel = ins if uniform else [int(v) for v in rng.choice(ins, size=int(rng.integers(1, n_in + 1)), replace=False)]
            tabs[o] = (a.mux_table([(i, float(rng.uniform(-0.7, 0.7))) for i in sel]),
                       a.delay_param(4000, int(rng.integers(0, 3900)), fs) if rng.random() < 0.6 else None)
        for o in outs:
            mux, dl = tabs[o]
            r = rng.random()
            if uniform or r < 0.5:
                a.load_mux(mux)
            elif r < 0.8:
                a.load_gain(int(rng.choice(ins)), float(rng.uniform(0.2, 1.0)))
            elif r < 0.9:
                a.load(int(rng.choice(ins)))
            else:
                a.load_store([(int(rng.choice(ins)), o)])     # raw copy: bypasses the ALU and the STORE mask
                continue
            if not uniform and rng.random() < 0.2:
                a.gain(float(rng.uniform(0.4, 1.2)))
            f = fin if uniform else int(rng.integers(0, 5))
            if f == 0:
                a.sat0db()
            elif f == 1:
                a.sat0db_tpdf()
            elif f == 2:
                a.sat0db_gain(float(rng.uniform(0.4, 1.0)))
            elif f == 3:
                a.sat0db_tpdf_gain(float(rng.uniform(0.4, 1.0)))
            if dl is not None:
                a.delay(dl)
            a.store(o)
    return a.end()


def _check(oracle_lib, w, fmt, fs, rng, what):
    S = int(rng.choice([1, 2, 5, 40, 150]))
    T = int(rng.choice([1, 3, 255, 256, 257, 1000, 5000 if S <= 5 else 700]))
    cut = sorted(set(int(v) for v in rng.integers(0, T + 1, size=2)))
    seeds = np.arange(S, dtype=np.int32) * 5 + 2
    ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    x = gen(str(rng.choice(["full", "noise", "impulse"])), S, T, max(ex.n_in, 1), fs)[:, :, : ex.n_in]
    ys, sts = oracle_lib.run_streams(w, fmt, fs, x, seeds=seeds, dither=24)
    planar = bool(rng.random() < 0.3)
    parts = []
    for a_, b_ in zip([0] + cut, cut + [T]):
        if b_ > a_:
            if planar:
                xp = np.ascontiguousarray(x[:, a_:b_].transpose(0, 2, 1))
                parts.append(ex.process(xp, layout=PLANAR).transpose(0, 2, 1))
            else:
                parts.append(ex.process(x[:, a_:b_]))
    y = np.concatenate(parts, axis=1)
    info = f"{what} kernel {ex.last_kernel} S={S} T={T} cut={cut} planar={planar}\n" + "\n".join(wire.disassemble(w))
    assert np.array_equal(y, ys), f"{np.count_nonzero(y != ys)} samples differ, channels {sorted(set(np.nonzero(y != ys)[2]))}, first frame {np.nonzero(y != ys)[1].min()}: " + info
    for s_ in sorted({0, S // 2, S - 1}):
        got, exp = ex.get_state(s_), expected_state(ex, sts[s_])
        assert np.array_equal(got, exp), f"state of stream {s_} differs at {np.nonzero(got != exp)[0][:8]}: " + info
    return ex.last_kernel


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("seed", range(12))
def test_random_mix_programs(oracle_lib, seed, fmt):
    rng = np.random.default_rng(7000 + seed)
    hits = 0
    for k in range(5):
        w = random_mix_program(rng, 48000, fmt)
        hits += _check(oracle_lib, w, fmt, 48000, rng, f"fmt {fmt} seed {seed}/{k}") == "mix"
    if fmt == 2:
        test_random_mix_programs.hits = getattr(test_random_mix_programs, "hits", 0) + hits


def test_mix_fuzz_reaches_the_mix_kernel():
    assert getattr(test_random_mix_programs, "hits", 0) >= 30, getattr(test_random_mix_programs, "hits", 0)


def random_fir_program(rng, fs=48000, fmt=2):
    a = wire.Asm(fmt=fmt, fmin=fs, fmax=fs)
    nch = int(rng.choice([1, 2, 3]))
    for k in range(nch):
        a.core()
        a.param()
        n = int(rng.choice([1, 2, 7, 64, 300, 1024]))
        r = np.random.default_rng(int(rng.integers(1 << 30)))
        c = r.standard_normal(n) * np.exp(-np.arange(n) / max(n / 5.0, 1.0))
        c *= 0.8 / np.abs(c).sum()
        taps = [wire.q28(float(v)) for v in c] if fmt == 2 else [float(np.float32(v)) for v in c]
        where = a.fir_impulses([taps])
        if rng.random() < 0.7:
            a.load_gain(8 + int(rng.integers(0, 2)), float(rng.uniform(0.3, 1.0)))
        else:
            a.load(8 + int(rng.integers(0, 2)))
        a.fir(where, n)
        r2 = rng.random()
        if r2 < 0.2:
            a.gain(float(rng.uniform(0.5, 1.2))); a.sat0db()
        elif r2 < 0.4:
            a.sat0db_gain(float(rng.uniform(0.5, 1.0)))
        else:
            a.sat0db()
        a.store(k)
    return a.end()


@pytest.mark.parametrize("fmt", [2, 3])
@pytest.mark.parametrize("seed", range(8))
def test_random_fir_programs(oracle_lib, seed, fmt):
    rng = np.random.default_rng(8000 + seed)
    for k in range(4):
        w = random_fir_program(rng, 48000, fmt)
        _check(oracle_lib, w, fmt, 48000, rng, f"fmt {fmt} seed {seed}/{k}")
