import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; if someone runs the whole suite on a CPU box, skip them.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def program_path(name):
    for ext in (".bin", ".h"):
        p = os.path.join(GOLDEN, "programs", name + ext)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(name)


def load_program(name):
    from avdsp_b200 import program
    return program.load(program_path(name))


def vector_names():
    d = os.path.join(GOLDEN, "vectors")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz"))


def load_vector(name):
    z = np.load(os.path.join(GOLDEN, "vectors", name + ".npz"))
    fmt, fs, seed, dither, frames = (int(v) for v in z["meta"])
    return dict(x=z["x"], y=z["y"], data=z["data"], code=z["code"], fmt=fmt, fs=fs, seed=seed, dither=dither,
                frames=frames, program=str(z["program"]), stimulus=str(z["stimulus"]))


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


# (program, DSP_FORMAT, fs) cases shared by several test modules
CASES = [
    ("c1_crossover2x2lfe_f2_48k", 2, 48000),
    ("c2_testrpi_xover_f2_192k", 2, 192000),
    ("c2_testrpi_xover_f2_multifs", 2, 88200),
    ("c3_peq16_f2_48k", 2, 48000),
    ("c3_peq16_f3_48k", 3, 48000),
    ("c3_peq16_f4_48k", 4, 48000),
    ("c3_peq16_f5_48k", 5, 48000),
    ("c3_peq16_f6_48k", 6, 48000),
    ("c5_mixer8x8_f2_192k", 2, 192000),
    ("ref_crossoverLV6", 2, 96000),
    ("ref_dacdiy1", 2, 192000),
    ("ref_dsptest1", 3, 48000),
    ("ref_dac8prodsp", 2, 96000),
    # DSP_FIR (C4): 4096-tap room-correction paths, and a small multi-rate program (convolution / plain delay / skipped)
    ("c4_fir4096_f2_48k", 2, 48000),
    ("c4_fir4096_f3_48k", 3, 48000),
    ("c4s_fir_f2_multifs", 2, 48000),
    ("c4s_fir_f2_multifs", 2, 96000),
    ("c4s_fir_f3_multifs", 3, 48000),
    ("c4s_fir_f3_multifs", 3, 96000),
    ("c4s_fir_f4_multifs", 4, 48000),
    ("c4s_fir_f5_multifs", 5, 48000),
    ("c4s_fir_f6_multifs", 6, 88200),
    # every opcode of the runtime's switch, every DSP_FORMAT (oracle/progs/allops_*.c, make_golden.asm_misc_program)
    *[(f"allops_alu_f{f}_48k", f, 48000) for f in (2, 3, 4, 5, 6)],
    *[(f"allops_gen_f{f}_multifs", f, fs) for f, fs in ((2, 44100), (2, 192000), (3, 88200), (4, 176400), (5, 48000), (6, 96000))],
    *[(f"allops_misc_f{f}_multifs", f, fs) for f, fs in ((2, 96000), (3, 48000), (4, 88200), (5, 44100), (6, 96000))],
    # the remaining checked-in fixtures of the reference (X/Y crossovers, DELAY_DP, SHIFT, SAT0DB_GAIN, MEM hand-offs)
    ("ref_dacfabriceo", 2, 96000),
    ("ref_dacfabriceo_oppo", 2, 44100),
    ("ref_lxmini_lr2", 2, 96000),
    ("ref_lxmini_lv8", 2, 192000),
    ("ref_win_mydspcode", 2, 88200),
]
