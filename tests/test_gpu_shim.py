"""The ALSA plugin rebuilt on the batched call (shim/avdsp_plugin_b200.c), driven period by period by shim/shim_harness the
way alsa-lib drives linux/avdsp_plugin.c (open with asound.conf keys, init at the stream rate, one transfer per period),
against the oracle's restatement of the OLD plugin loop (`process_plugin_order`, linux/avdsp_plugin.c:95-142) and against
the canonical order.  S16 / S24_3LE / S32 input, S32 output."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_program, program_path
from avdsp_b200 import synth
from oracle import wire

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "shim", "shim_harness")


def encode(x32, fmt):
    """int32 s.31 frames -> the bytes ALSA would hand over, and the s.31 values the plugin widens them to"""
    if fmt == "s32":
        return x32.astype("<i4").tobytes(), x32
    if fmt == "s16":
        x16 = (x32 >> 16).astype("<i2")
        return x16.tobytes(), x16.astype(np.int32) << 16
    x24 = x32 & ~0xFF
    b = np.empty(x24.shape + (3,), np.uint8)
    b[..., 0] = (x24 >> 8) & 0xFF; b[..., 1] = (x24 >> 16) & 0xFF; b[..., 2] = (x24 >> 24) & 0xFF
    return b.tobytes(), x24


def run_harness(tmp_path, prog_file, rate, fmt, period, raw, n_out, *keys):
    if not os.path.exists(HARNESS):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "shim")])
    fi, fo = tmp_path / "in.raw", tmp_path / "out.raw"
    fi.write_bytes(raw)
    r = subprocess.run([HARNESS, str(prog_file), str(rate), fmt, str(period), str(fi), str(fo), *keys], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.frombuffer(fo.read_bytes(), dtype="<i4").reshape(-1, n_out), r.stdout


def channels(w):
    ins, outs = wire.io_maps(w)
    ins, outs = [i for i in ins if i < 16], [o for o in outs if o < 16]
    return ins, outs, max(ins) - 8 + 1, max(outs) + 1


def _talking_cores_program():
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.tpdf_calc(20); a.load_gain(8, 0.5); a.sat0db_tpdf(); a.store(0)
    a.core(); a.load_gain(9, 0.25); a.sat0db_tpdf(); a.store(1); a.load(8); a.store(2)
    return a.end()


def _sparse_io_program():
    """skips io slots on both sides: PCM has 3 input and 4 output channels, the program uses in 8, 10 and out 1, 3"""
    a = wire.Asm(fmt=2, fmin=48000, fmax=48000)
    a.core(); a.load_gain(8, 0.5); a.sat0db(); a.store(1)
    a.core(); a.load_gain(10, 0.25); a.sat0db(); a.store(3)
    return a.end()


@pytest.mark.parametrize("fmt", ["s16", "s24", "s32"])
@pytest.mark.parametrize("which", ["c2", "talking", "sparse"])
def test_shim_equals_the_old_plugin_loop(oracle_lib, tmp_path, which, fmt):
    if which == "c2":
        w, rate, path = load_program("c2_testrpi_xover_f2_192k"), 192000, program_path("c2_testrpi_xover_f2_192k")
    else:
        w, rate = (_talking_cores_program() if which == "talking" else _sparse_io_program()), 48000
        path = tmp_path / "prog.bin"
        w.astype("<i4").tofile(path)
    ins, outs, cin, cout = channels(w)
    T, period = 700, 128                                   # ragged last period
    x = np.zeros((T, cin), np.int32)
    x[:, [i - 8 for i in ins]] = synth.pcm("full", 1, T, len(ins), rate)[0]
    raw, xw = encode(x, fmt)
    y, log = run_harness(tmp_path, path, rate, fmt, period, raw, cout, "dither=24")
    assert f"nbchanin {cin}, nbchanout {cout}" in log
    o = oracle_lib.Oracle(w, 2, rate, seed=0, dither=24)
    yo = o.process_plugin_order(xw, period, cin, cout)
    assert np.array_equal(y[:, outs], yo[:, outs]), "shim (order plugin) differs from the old plugin's loop nest"
    # canonical order on request
    yc, _ = run_harness(tmp_path, path, rate, fmt, period, raw, cout, "dither=24", "order=canonical")
    oc = oracle_lib.Oracle(w, 2, rate, seed=0, dither=24)
    assert np.array_equal(yc[:, outs], oc.process(xw[:, [i - 8 for i in ins]])[:, [outs.index(k) for k in outs]])
    if which == "talking":
        assert not np.array_equal(y, yc), "the program must tell the two orders apart"


def test_shim_tagoutput_and_rate_check(oracle_lib, tmp_path):
    w, rate = load_program("c2_testrpi_xover_f2_192k"), 192000
    path = program_path("c2_testrpi_xover_f2_192k")
    ins, outs, cin, cout = channels(w)
    T, period = 300, 64
    x = synth.pcm("noise", 1, T, cin, rate)[0]
    y, _ = run_harness(tmp_path, path, rate, "s32", period, x.astype("<i4").tobytes(), cout, "tagoutput=24")
    o = oracle_lib.Oracle(w, 2, rate, seed=0, dither=31)
    yo = o.process_plugin_order(x, period, cin, cout)
    # linux/avdsp_plugin.c:132-137 in the old loop order: per period, core-major, the first output of every core
    firsts = []
    p = 0
    for at, op, sk in wire.walk(w.view(np.uint32)):
        if op == wire.OP["CORE"]:
            om = int(w[at + 2]) & 0xFFFF
            firsts.append(min(k for k in range(16) if om >> k & 1))
    prev = 0
    for b0 in range(0, T, period):
        for ch in firsts:
            for n in range(b0, min(b0 + period, T)):
                top = int(yo[n, ch]) & ~0xFFFF
                top = top - (1 << 32) if top >= 1 << 31 else top
                yo[n, ch] = np.int32(top | (prev & 0xFF00))
                prev = (top >> 8) + 0x100
    assert np.array_equal(y, yo)
    # a rate outside the program's range is refused at init like dspRuntimeReset does (-2 -> EINVAL)
    r = subprocess.run([HARNESS, str(path), "48000", "s32", "64", str(tmp_path / "in.raw"), str(tmp_path / "o.raw")], capture_output=True, text=True)
    assert r.returncode != 0 and "init failed" in r.stderr
