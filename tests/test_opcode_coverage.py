"""Every opcode of the runtime's switch (runtime/dsp_runtime.c:319-1305) is executed by at least one golden vector that the
compiled reference produced (tests/golden/make_golden.py) -- which the oracle is pinned to on the CPU
(test_oracle_golden.py) and the CUDA kernels on the GPU (test_gpu_parity.py::test_golden_vectors_*).

`python tests/test_opcode_coverage.py` prints the "opcode -> golden that executes it" table of DESIGN.md."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import load_program, load_vector, vector_names   # noqa: E402
from oracle import wire                                          # noqa: E402

# opcodes that have no `case` doing work: structure words, skipped at run time (:321-335, :852-868)
STRUCTURAL = {"END_OF_CODE", "HEADER", "NOP", "CORE", "PARAM", "PARAM_NUM"}


# first line of each opcode's `case` in /root/reference/module_avdsp/runtime/dsp_runtime.c
RT_LINE = {"SWAPXY": 337, "COPYXY": 344, "COPYYX": 349, "CLRXY": 354, "ADDXY": 360, "ADDYX": 365, "SUBXY": 370, "SUBYX": 375, "NEGX": 380,
           "NEGY": 385, "SHIFT": 390, "MULXY": 408, "DIVXY": 414, "DIVYX": 420, "AVGXY": 426, "AVGYX": 432, "SQRTX": 438, "SAT0DB": 464,
           "SAT0DB_TPDF": 478, "SAT0DB_GAIN": 494, "SAT0DB_TPDF_GAIN": 514, "TPDF_CALC": 537, "TPDF": 547, "LOAD": 565, "LOAD_GAIN": 586,
           "STORE": 610, "GAIN": 636, "VALUE": 643, "VALUE_INT": 651, "WHITE": 664, "MUL_VALUE": 678, "DIV_VALUE": 685, "MUL_VALUE_INT": 692,
           "DIV_VALUE_INT": 703, "AND_VALUE_INT": 714, "DELAY_1": 726, "LOAD_STORE": 738, "LOAD_MEM": 750, "STORE_MEM": 760, "DELAY": 769,
           "DELAY_DP": 798, "BIQUADS": 827, "SERIAL": 863, "LOAD_MUX": 871, "DATA_TABLE": 900, "FIR": 928, "FIR(delay)": 940, "RMS": 972,
           "DCBLOCK": 1063, "DITHER": 1112, "DITHER_NS2": 1138, "DISTRIB": 1175, "DIRAC": 1213, "SQUAREWAVE": 1234, "CLIP": 1264,
           "LOAD_MEM_DATA": 1277, "SINE": 1284}


def executed_opcodes(words, fs):
    """Opcode names inside the cores of the program (what dspRuntime walks), minus forms that are skipped at this fs."""
    w = np.asarray(words).view(np.uint32)
    h = wire.header(w)
    fi = wire.freq_index(fs) - h["freqMin"]
    nf = h["freqMax"] - h["freqMin"] + 1
    names = set()
    in_core = h["numCores"] <= 1 and not any(op == wire.OP["CORE"] for _, op, _ in wire.walk(w))
    for p, op, skip in wire.walk(w):
        if skip == 0:
            break
        name = wire.OPCODES[op]
        if name == "CORE":
            in_core = True
        if not in_core or name in ("PARAM", "PARAM_NUM"):
            names.add(name) if name in STRUCTURAL else None
            continue
        if name == "FIR":
            rel = int(np.int32(w[p + 1 + fi]))
            if rel == 0:
                continue
            lw = int(w[p + rel])
            names.add("FIR(delay)" if lw >> 16 else "FIR")
            continue
        names.add(name)
    return names


def coverage():
    table = {}
    for v in vector_names():
        vec = load_vector(v)
        for name in executed_opcodes(load_program(vec["program"]), vec["fs"]):
            table.setdefault(name, []).append((v, vec["fmt"]))
    return table


def test_every_opcode_is_executed_by_a_reference_made_golden():
    table = coverage()
    missing = [n for n in wire.OPCODES if n not in STRUCTURAL and n not in table]
    assert not missing, f"no golden vector executes: {missing}"
    assert "FIR(delay)" in table
    # every opcode with format-dependent arithmetic is pinned in every DSP_FORMAT
    for name in wire.OPCODES:
        if name in STRUCTURAL or name in ("SINE", "SERIAL", "LOAD_MUX", "LOAD_MEM", "STORE_MEM", "LOAD_STORE", "FIR"):
            continue
        fmts = {f for _, f in table[name]}
        assert fmts >= {2, 3, 4, 5, 6}, (name, sorted(fmts))
    assert {f for _, f in table["FIR"]} >= {3, 4, 5, 6}       # fixed-point FIR: the reference kernel is broken


if __name__ == "__main__":
    import re
    t = coverage()
    print("| opcode | `dsp_runtime.c` | formats pinned | golden vectors that execute it (families; `tests/golden/vectors/*.npz`) |")
    print("|---|---|---|---|")
    for n in wire.OPCODES + ["FIR(delay)"]:
        if n in STRUCTURAL:
            continue
        vs = t.get(n, [])
        fm = ",".join(str(f) for f in sorted({f for _, f in vs}))
        fam = sorted({re.sub(r"_f\d.*$|_\d+k.*$|_(noise|full|sine|impulse)$", "", v) for v, _ in vs})
        print(f"| {n} | :{RT_LINE.get(n, '')} | {fm} | {', '.join(fam)} ({len(vs)} vectors) |")
