"""Randomised CHAIN-shaped programs (source -> biquad cascade -> finish -> delay -> stores, several per core) at random batch
shapes through every chain kernel that accepts them -- k_chain2, k_chain3 (forced, several streams per CTA), AUTO -- against the
oracle, bit for bit in outputs and every state word: fixed point, float format 3 and float format 5 (float samples).  The point is the geometry: part cuts of odd cascade lengths, lags
longer than a call, calls shorter than a tile, partial CTAs, delay rings longer and shorter than the row ring."""
import numpy as np
import pytest

from avdsp_b200 import Executor, AvdspError, synth, KERNEL_AUTO, KERNEL_GENERIC, KERNEL_CHAIN_V2, KERNEL_CHAIN_V3
from oracle import wire
from test_gpu_parity import expected_state
from test_gpu_chain3 import _float_state_close

pytestmark = pytest.mark.gpu


def random_chain_program(rng, fs=48000, fmt=2, unstable=False):
    a = wire.Asm(fmt=fmt, fmin=fs, fmax=fs)
    outs = list(range(8))
    rng.shuffle(outs)
    plain = rng.random() < 0.4                           # only LOAD / LOAD_GAIN sources and plain finishes: the shape k_chain3 takes
    first = True
    for c in range(int(rng.integers(1, 4))):
        a.core()
        if first and rng.random() < 0.8:
            a.tpdf_calc(int(rng.choice([16, 20, 24, 31])))
        first = False
        a.param()
        for _path in range(int(rng.integers(1, 4))):
            if not outs:
                break
            nsec = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 16]))
            sec = a.biquad_sections([[wire.rbj_peak(fs, float(rng.uniform(60, 15000)), float(rng.uniform(-4, 4) if unstable else rng.uniform(0.5, 4)), float(rng.uniform(0.5, 1.8)))]
                                     for _ in range(nsec)])
            dl = a.delay_param(3000, int(rng.integers(0, 2900)), fs) if rng.random() < 0.5 else None
            mux = a.mux_table([(8, float(rng.uniform(-0.6, 0.6))), (9, float(rng.uniform(-0.6, 0.6)))])
            r = rng.random()
            if r < 0.3:
                a.load(int(rng.choice([8, 9])))
            elif r < 0.85 or plain:
                a.load_gain(int(rng.choice([8, 9])), float(rng.choice([1.0, float(rng.uniform(0.2, 1.0))])))
            else:
                a.load_mux(mux)
            if not plain and dl is not None and rng.random() < 0.2:
                a.delay(dl); dl = None                     # delay in front of the cascade
            a.biquads(sec)
            if not plain and rng.random() < 0.25:
                a.gain(float(rng.uniform(0.4, 1.1)))
            r = rng.random()
            if r < 0.4:
                a.sat0db()
            elif r < 0.8:
                a.sat0db_tpdf()
            elif plain:
                a.sat0db()
            elif r < 0.9:
                a.sat0db_gain(float(rng.uniform(0.4, 1.0)))
            else:
                a.sat0db_tpdf_gain(float(rng.uniform(0.4, 1.0)))
            if dl is not None:
                a.delay(dl)
            a.store(outs.pop())
            if outs and rng.random() < 0.2:
                a.store(outs.pop())
    return a.end()


@pytest.mark.parametrize("fmt", [2, 3, 5])
@pytest.mark.parametrize("seed", range(16))
def test_random_chain_programs(oracle_lib, monkeypatch, seed, fmt):
    rng = np.random.default_rng(5000 + seed)
    fs = 48000
    ran = {"chain2": 0, "chain3": 0}
    for k in range(4):
        # seeds 12..15: some sections get a negative Q and blow up -- wrapping / saturation replay in fixed point, infinities and NaNs in
        # the float formats (the exact second pass of the float class: dspMulFloatFloat's integer form, the host's NaN rules)
        w = random_chain_program(rng, fs, fmt, unstable=seed >= 12)
        S = int(rng.choice([1, 3, 7, 33, 70]))
        T = int(rng.choice([1, 2, 31, 97, 333, 1700]))
        cut = sorted(set(int(v) for v in rng.integers(0, T + 1, size=2)))
        seeds = np.arange(S, dtype=np.int32) * 3 + seed
        x = (synth.pcm_float if fmt >= 5 else synth.pcm)(str(rng.choice(["full", "noise", "impulse"])), S, T, 2, fs)
        monkeypatch.setenv("AVDSP_B200_NS3", str(int(rng.choice([0, 5, 32]))))
        monkeypatch.setenv("AVDSP_B200_PART3", str(int(rng.choice([4, 4, 8, 2]))))
        ys = sts = None
        for kern in (KERNEL_AUTO, KERNEL_CHAIN_V2, KERNEL_CHAIN_V3):
            ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
            ex.set_kernel(kern)
            xi = x[:, :, : ex.n_in]
            if ys is None:
                ys, sts = oracle_lib.run_streams(w, fmt, fs, xi, seeds=seeds, dither=24)
            try:
                parts = [ex.process(xi[:, a_:b_]) for a_, b_ in zip([0] + cut, cut + [T]) if b_ > a_]
            except AvdspError:
                assert kern != KERNEL_AUTO                 # a forced chain kernel may decline a shape; AUTO never fails
                continue
            y = np.concatenate(parts, axis=1)
            name = ex.last_kernel + (str(ex.last_chain_variant) if ex.last_kernel == "chain" else "")
            if name in ran:
                ran[name] += 1
            what = f"fmt {fmt} seed {seed}/{k} kernel {name} S={S} T={T} cut={cut} NS3/PART3 env\n" + "\n".join(wire.disassemble(w))
            assert np.array_equal(y, ys), f"{np.count_nonzero(y != ys)} samples differ, channels {sorted(set(np.nonzero(y != ys)[2]))}: " + what
            for s_ in sorted({0, S // 2, S - 1}):
                got, exp = ex.get_state(s_), expected_state(ex, sts[s_])
                if fmt == 2 or ex.last_kernel == "generic":
                    assert np.array_equal(got, exp), f"state of stream {s_} differs at {np.nonzero(got != exp)[0][:8]}: " + what
                else:
                    from test_gpu_fuzz import nan_aware_equal      # two NaNs are equal whatever their payload (see there)
                    assert nan_aware_equal(got, exp), f"float state of stream {s_} differs at {np.nonzero(got != exp)[0][:8]}: " + what
    test_random_chain_programs.ran = {n: getattr(test_random_chain_programs, "ran", {}).get(n, 0) + v for n, v in ran.items()}


def test_chain_fuzz_reaches_both_chain_kernels():
    ran = getattr(test_random_chain_programs, "ran", {})
    assert ran.get("chain2", 0) >= 60 and ran.get("chain3", 0) >= 20, ran
