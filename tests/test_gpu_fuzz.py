"""Randomised X/Y-dataflow programs (wire assembler) through whatever kernel AUTO picks -- k_dag where the decoder's symbolic
X/Y execution recognises the program, the interpreter otherwise -- against the oracle, outputs and state, bit for bit.
The generator only emits sequences that are meaningful in the reference (no division, no store of an undefined ALU), but
it does not know what the DAG recognition accepts: both outcomes are checked, and the test requires that a fair share of
the programs does land on k_dag."""
import numpy as np
import pytest

from avdsp_b200 import Executor, synth, KERNEL_GENERIC
from oracle import wire

pytestmark = pytest.mark.gpu


def random_program(rng, fs=48000, fmt=2):
    a = wire.Asm(fmt=fmt, fmin=fs, fmax=fs)
    n_cores = int(rng.integers(1, 4))
    out_slots = list(range(0, 8))
    rng.shuffle(out_slots)
    mems = []                                           # (address, defined)
    first = True
    for c in range(n_cores):
        a.core()
        if first and rng.random() < 0.7:
            a.tpdf_calc(int(rng.choice([20, 24, 31])))
        first = False
        a.param()
        secs = [a.biquad_sections([[wire.rbj_peak(fs, float(rng.uniform(60, 15000)), float(rng.uniform(0.5, 4)), float(rng.uniform(0.5, 1.8)))]
                                   for _ in range(int(rng.integers(1, 6)))]) for _ in range(3)]
        dls = [a.delay_param(3000, int(rng.integers(20, 2800)), fs) for _ in range(3)]
        new_mems = [a.mem_location() for _ in range(2)]
        x_def = y_def = False                            # X / Y hold an expression the reference computes meaningfully
        x_fin = False                                    # X is a saturated sample
        for _path in range(int(rng.integers(1, 4))):
            # a source
            r = rng.random()
            if r < 0.35:
                a.load(int(rng.choice([8, 9])))
            elif r < 0.8 or not any(d for _, d in mems):
                a.load_gain(int(rng.choice([8, 9])), float(rng.uniform(0.2, 1.0)))
            else:
                a.load_mem(rng.choice([m for m, d in mems if d]))
            y_def, x_def, x_fin = x_def and not x_fin, True, False
            fresh = False
            for _step in range(int(rng.integers(1, 8))):
                r = rng.random()
                if r < 0.12:
                    a.simple("COPYXY"); y_def = x_def
                elif r < 0.22 and y_def:
                    a.simple("SWAPXY"); fresh = False
                elif r < 0.30 and y_def and x_def:
                    a.simple(str(rng.choice(["ADDXY", "SUBXY", "SUBYX", "ADDYX"]))); fresh = False
                elif r < 0.55:
                    a.biquads(secs[int(rng.integers(0, 3))]); fresh = True
                elif r < 0.63:
                    a.gain(float(rng.uniform(0.3, 1.2)))
                elif r < 0.68:
                    a.shift(int(rng.choice([-100, -28, -3])))
                elif r < 0.74:
                    a.delay(dls[int(rng.integers(0, 3))], dp=bool(rng.random() < 0.5))
                elif r < 0.80 and fresh and new_mems:
                    m = new_mems.pop(); a.store_mem(m); mems.append((m, True))
                elif r < 0.84:
                    a.simple("CLRXY"); y_def = False; fresh = False
                    a.load_gain(int(rng.choice([8, 9])), 0.5)
            # a finish and one or two stores
            r = rng.random()
            if r < 0.3:
                a.sat0db()
            elif r < 0.55:
                a.sat0db_tpdf()
            elif r < 0.75:
                a.sat0db_gain(float(rng.uniform(0.4, 1.0)))
            elif r < 0.9:
                a.sat0db_tpdf_gain(float(rng.uniform(0.4, 1.0)))
            # else: DSP_STORE of the unsaturated value
            x_fin = True
            if out_slots:
                a.store(out_slots.pop())
            if rng.random() < 0.4 and out_slots:
                a.delay(dls[int(rng.integers(0, 3))])
                a.store(out_slots.pop())
            if rng.random() < 0.5 and y_def:
                a.simple("SWAPXY"); x_fin = False; x_def = True; y_def = False
                if rng.random() < 0.5:
                    a.biquads(secs[int(rng.integers(0, 3))])
                a.sat0db()
                if out_slots:
                    a.store(out_slots.pop())
                x_fin = True
        if not out_slots:
            break
    return a.end()


def random_misc_program(rng, fs=48000, fmt=2):
    """The opcodes the X/Y generator above never emits: immediates, products and quotients of the registers, LOAD_MUX, LOAD_STORE,
    core-local TPDF tables, WHITE, DITHER, DITHER_NS2, DCBLOCK, CLIP, DIRAC / SQUAREWAVE, DELAY_1 and the fixed delays.
    Divisors are non-zero constants (the reference divides by whatever Y holds)."""
    a = wire.Asm(fmt=fmt, fmin=fs, fmax=fs)
    outs = list(range(8))
    rng.shuffle(outs)
    for c in range(int(rng.integers(1, 4))):
        a.core()
        if c == 0:
            a.tpdf_calc(int(rng.choice([16, 20, 24])))
        elif rng.random() < 0.5:
            a.tpdf(int(rng.choice([0, 16, 20, 24])))                 # core-local table: also changes this core's STORE mask
        a.param()
        sec = a.biquad_sections([[wire.rbj_peak(fs, float(rng.uniform(80, 12000)), float(rng.uniform(0.5, 3)), float(rng.uniform(0.5, 1.5)))]
                                 for _ in range(int(rng.integers(1, 4)))])
        mux = a.mux_table([(8, float(rng.uniform(-0.6, 0.6))), (9, float(rng.uniform(-0.6, 0.6)))])
        ns2 = a.num(float(rng.uniform(1.2, 2.2))); a.num(float(rng.uniform(-1.6, -0.6))); a.num(float(rng.uniform(0.1, 0.5)))
        for _path in range(int(rng.integers(1, 4))):
            if not outs:
                break
            r = rng.random()
            if r < 0.25:
                a.load(int(rng.choice([8, 9])))
            elif r < 0.55:
                a.load_gain(int(rng.choice([8, 9])), float(rng.uniform(0.2, 1.0)))
            elif r < 0.7:
                a.load_mux(mux)
            elif r < 0.8:
                a.simple("CLRXY"); a.dirac(float(rng.uniform(0.2, 0.9)), [int(rng.integers(3, 40))], square=bool(rng.random() < 0.5))
            elif r < 0.9:
                a.simple("WHITE")
            else:
                a.value(float(rng.uniform(-0.5, 0.5)))
            for _step in range(int(rng.integers(0, 6))):
                r = rng.random()
                if r < 0.10:
                    a.gain(float(rng.uniform(0.3, 1.2)))
                elif r < 0.18:
                    a.imm("MUL_VALUE", float(rng.uniform(0.3, 1.5)))
                elif r < 0.26:
                    a.imm("DIV_VALUE", float(rng.choice([-2.0, 0.75, 1.5, 3.0])))
                elif r < 0.32:
                    a.imm("MUL_VALUE_INT", int(rng.choice([2, 3, -2])), as_int=True)
                elif r < 0.38:
                    a.imm("DIV_VALUE_INT", int(rng.choice([2, 3, -5])), as_int=True)
                elif r < 0.44:
                    a.value_int(int(rng.choice([2, 3, 5]))); a.simple("SWAPXY"); a.simple(str(rng.choice(["DIVXY", "MULXY"])))
                elif r < 0.50:
                    a.simple("COPYXY"); a.simple(str(rng.choice(["AVGXY", "AVGYX", "NEGX", "NEGY", "ADDXY"])))
                elif r < 0.56:
                    a.shift(int(rng.choice([-2, -1, 1])))
                elif r < 0.62:
                    a.delay_1()
                elif r < 0.70:
                    a.delay_fixed_us(int(rng.integers(50, 900)), fs, dp=bool(rng.random() < 0.5))
                elif r < 0.78:
                    a.dcblock([float(rng.choice([-0.003, -0.0015, -0.0008]))])
                elif r < 0.84:
                    a.clip(float(rng.uniform(0.05, 0.6)))
                elif r < 0.94:
                    a.biquads(sec)
            r = rng.random()
            if r < 0.15:
                a.dither(); a.sat0db()
            elif r < 0.30:
                a.dither_ns2(ns2); a.sat0db()
            elif r < 0.45:
                a.sat0db()
            elif r < 0.6:
                a.sat0db_tpdf()
            elif r < 0.75:
                a.sat0db_gain(float(rng.uniform(0.4, 1.0)))
            elif r < 0.9:
                a.sat0db_tpdf_gain(float(rng.uniform(0.4, 1.0)))
            a.store(outs.pop())
        if outs and rng.random() < 0.5:
            a.load_store([(int(rng.choice([8, 9])), outs.pop())])
    return a.end()


@pytest.mark.parametrize("seed", range(24))
def test_random_xy_programs(oracle_lib, seed):
    rng = np.random.default_rng(1000 + seed)
    fs = 48000
    hits = 0
    for k in range(6):
        w = random_program(rng, fs)
        S, T = (5, 210) if k < 3 else (int(rng.choice([1, 33, 70])), int(rng.choice([1, 2, 31, 333, 1000])))
        cut = 77 if k < 3 else int(rng.integers(0, T + 1))
        seeds = np.arange(S, dtype=np.int32) + seed
        try:
            ex = Executor(w, fs, 2, S, seeds=seeds, dither=24)
        except Exception:
            continue                                      # (a generated program the decoder refuses: not this test's subject)
        x = synth.pcm("full" if k & 1 else "noise", S, T, ex.n_in, fs) if ex.n_in else np.zeros((S, T, 0), np.int32)
        ys, sts = oracle_lib.run_streams(w, 2, fs, x, seeds=seeds, dither=24)
        y = np.concatenate([ex.process(x[:, a_:b_]) for a_, b_ in ((0, cut), (cut, T)) if b_ > a_], axis=1)
        kern = ex.last_kernel
        hits += kern == "dag"
        assert np.array_equal(y, ys), f"seed {seed}/{k} [{kern}] S={S} T={T} cut={cut}: {np.count_nonzero(y != ys)} samples differ\n" + "\n".join(wire.disassemble(w))
        for s in (0, S - 1):
            data, aux, code = sts[s]
            st = ex.get_state(s)
            assert np.array_equal(st[: ex.data_size], data), f"seed {seed}/{k} [{kern}]: data area differs at {np.nonzero(st[:ex.data_size] != data)[0][:8]}\n" + "\n".join(wire.disassemble(w))
            for q, wd in enumerate(ex.mem_words):
                assert np.array_equal(st[ex.mem_offset + 2 * q: ex.mem_offset + 2 * q + 2], code[wd: wd + 2]), f"seed {seed}/{k} [{kern}]: MEM word {q}"
    test_random_xy_programs.hits = getattr(test_random_xy_programs, "hits", 0) + hits


def nan_aware_equal(a, b, float_words=True):
    """bit-for-bit, except that two float NaNs are equal whatever their payload: an unstable random cascade ends in inf - inf,
    and which operand's payload an x86 addss keeps is the reference compiler's choice of operand order, not the reference's"""
    a, b = np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32)
    d = a != b
    if not d.any():
        return True
    if not float_words:
        return False
    isnan = lambda u: ((u & 0x7F800000) == 0x7F800000) & ((u & 0x007FFFFF) != 0)
    return bool(np.all(isnan(a[d]) & isnan(b[d])))


@pytest.mark.parametrize("fmt", [3, 4, 5, 6])
@pytest.mark.parametrize("seed", range(6))
def test_random_xy_programs_float_formats(oracle_lib, seed, fmt):
    """The same generator encoded for the float ALUs (DSP_FORMAT 3..6; 5/6 with float samples): the interpreter is the bit-exact
    path for those (outputs AND data area), AUTO's choice must produce the same outputs."""
    rng = np.random.default_rng(2000 + seed)
    fs = 48000
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    for k in range(6):
        w = random_program(rng, fs, fmt)
        S, T = 4, 150
        seeds = np.arange(S, dtype=np.int32) + seed
        ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
        ex.set_kernel(KERNEL_GENERIC)
        x = gen("full" if k & 1 else "noise", S, T, ex.n_in, fs)
        ys, sts = oracle_lib.run_streams(w, fmt, fs, x, seeds=seeds, dither=24)
        y = np.concatenate([ex.process(x[:, :61]), ex.process(x[:, 61:])], axis=1)
        assert nan_aware_equal(y, ys, fmt >= 5), f"fmt {fmt} seed {seed}/{k}: {np.count_nonzero(y != ys)} samples differ\n" + "\n".join(wire.disassemble(w))
        for s_ in (0, S - 1):
            data = sts[s_][0]
            st = ex.get_state(s_)
            assert nan_aware_equal(st[: ex.data_size], data), f"fmt {fmt} seed {seed}/{k}: data area differs at {np.nonzero(st[:ex.data_size] != data)[0][:8]}"
        ex2 = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
        y2 = np.concatenate([ex2.process(x[:, :61]), ex2.process(x[:, 61:])], axis=1)
        assert nan_aware_equal(y2, ys, fmt >= 5), f"fmt {fmt} seed {seed}/{k} [AUTO: {ex2.last_kernel}]: {np.count_nonzero(y2 != ys)} samples differ"


@pytest.mark.parametrize("fmt", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("seed", range(6))
def test_random_misc_programs(oracle_lib, seed, fmt):
    """random_misc_program in every format: interpreter bit-exact (outputs and data area), AUTO's choice the same outputs."""
    rng = np.random.default_rng(3000 + seed)
    fs = 48000
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    for k in range(6):
        w = random_misc_program(rng, fs, fmt)
        S, T = 4, 150
        seeds = np.arange(S, dtype=np.int32) + seed
        ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
        ex.set_kernel(KERNEL_GENERIC)
        x = gen("full" if k & 1 else "noise", S, T, max(ex.n_in, 1), fs)[:, :, : ex.n_in]
        ys, sts = oracle_lib.run_streams(w, fmt, fs, x, seeds=seeds, dither=24)
        y = np.concatenate([ex.process(x[:, :61]), ex.process(x[:, 61:])], axis=1)
        assert nan_aware_equal(y, ys, fmt >= 5), f"fmt {fmt} seed {seed}/{k}: {np.count_nonzero(y != ys)} samples differ, channels {sorted(set(np.nonzero(y != ys)[2]))}\n" + "\n".join(wire.disassemble(w))
        for s_ in (0, S - 1):
            data = sts[s_][0]
            st = ex.get_state(s_)
            assert nan_aware_equal(st[: ex.data_size], data, fmt != 2), f"fmt {fmt} seed {seed}/{k}: data area differs at {np.nonzero(st[:ex.data_size] != data)[0][:8]}\n" + "\n".join(wire.disassemble(w))
        ex2 = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
        y2 = np.concatenate([ex2.process(x[:, :61]), ex2.process(x[:, 61:])], axis=1)
        assert nan_aware_equal(y2, ys, fmt >= 5), f"fmt {fmt} seed {seed}/{k} [AUTO: {ex2.last_kernel}]: {np.count_nonzero(y2 != ys)} samples differ\n" + "\n".join(wire.disassemble(w))


@pytest.mark.parametrize("seed", range(12))
def test_random_programs_in_plugin_order(oracle_lib, seed):
    """The ALSA plugin's loop nest (core-major inside a period, linux/avdsp_plugin.c:95-142).  The decoder keeps a fused kernel
    only where it can prove that both orders give the same result (decoder.cpp analyseOrder); a wrong proof shows here: all three
    generators, random periods, two calls, against the oracle's restatement of the plugin loop."""
    from test_gpu_fuzz_chain import random_chain_program
    rng = np.random.default_rng(9000 + seed)
    fs = 48000
    fused = 0
    for k in range(6):
        w = (random_program, random_misc_program, random_chain_program)[k % 3](rng, fs, 2)
        S, T = 3, int(rng.choice([100, 300]))
        period = int(rng.choice([1, 16, 64, 128, 1000]))
        cut = int(rng.integers(1, T))
        seeds = np.arange(S, dtype=np.int32) + seed
        ex = Executor(w, fs, 2, S, seeds=seeds, dither=24)
        ex.set_order(period)
        x = synth.pcm("full" if k & 1 else "noise", S, T, max(ex.n_in, 1), fs)[:, :, : ex.n_in]
        y = np.concatenate([ex.process(x[:, :cut]), ex.process(x[:, cut:])], axis=1)
        fused += ex.last_kernel != "generic"
        nin = (max(ex.in_idx) - 8 + 1) if len(ex.in_idx) else 1
        nout = max(ex.out_idx) + 1
        for s_ in range(S):
            o = oracle_lib.Oracle(w, 2, fs, seed=int(seeds[s_]), dither=24)
            xin = np.zeros((T, nin), np.int32)
            for q, slot in enumerate(ex.in_idx):
                xin[:, slot - 8] = x[s_, :, q]
            yo = np.concatenate([o.process_plugin_order(xin[:cut], period, nin, nout), o.process_plugin_order(xin[cut:], period, nin, nout)])
            what = f"seed {seed}/{k} [{ex.last_kernel}] period {period} T={T} cut={cut}\n" + "\n".join(wire.disassemble(w)) + "\n" + ex.trace[-400:]
            assert np.array_equal(y[s_], yo[:, ex.out_idx]), f"stream {s_}: {np.count_nonzero(y[s_] != yo[:, ex.out_idx])} samples differ: " + what
            assert np.array_equal(ex.get_state(s_)[: ex.data_size], o.data), f"stream {s_}: data area differs: " + what
    test_random_programs_in_plugin_order.fused = getattr(test_random_programs_in_plugin_order, "fused", 0) + fused


def test_plugin_order_fuzz_keeps_fused_kernels_somewhere():
    assert getattr(test_random_programs_in_plugin_order, "fused", 0) >= 10, getattr(test_random_programs_in_plugin_order, "fused", 0)


def test_fuzz_reaches_the_dag_kernel():
    assert getattr(test_random_xy_programs, "hits", 0) >= 20, "the random programs hardly ever map to k_dag: the generator drifted"
