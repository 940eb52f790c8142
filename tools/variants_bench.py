"""per-stream parameter variants at batch width: python tools/variants_bench.py [streams] [frames]
C2-like crossover program (tests/test_gpu_params.crossover_program), V variants laid out in contiguous blocks or interleaved."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import json
import numpy as np, torch
from test_gpu_params import crossover_program, apply_diff
from avdsp_b200 import Executor, synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4800
fs = 48000
progs = [crossover_program(200.0 * 1.06 ** v, 100 + 25 * v, 0.5 + 0.005 * v, fs) for v in range(64)]
x = torch.from_numpy(synth.pcm("noise", S, T, 1, fs)).cuda()


def run(V, interleave):
    ex = Executor(progs[0], fs, 2, S, seeds=np.arange(S, dtype=np.int32), dither=24)
    y = torch.empty((S, T, ex.n_out), dtype=torch.int32, device="cuda")
    if V > 1:
        if interleave:                                  # stream s runs variant s % V: every run is one stream
            for s in range(S):
                if s % V: apply_diff(ex, s, 1, progs[0], progs[s % V])
        else:                                           # V contiguous blocks
            blk = S // V
            for v in range(1, V): apply_diff(ex, v * blk, blk if v < V - 1 else S - v * blk, progs[0], progs[v])
    for _ in range(2): ex.process(x, out=y)
    torch.cuda.synchronize()
    l0 = ex.launch_count
    t0 = time.perf_counter()
    for _ in range(3): ex.process(x, out=y)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps({"streams": S, "frames": T, "variants": ex.num_variants, "layout": "interleaved" if interleave else "blocks",
                      "launches_per_call": (ex.launch_count - l0) // 3, "kernel": ex.last_kernel + str(ex.last_chain_variant or ""),
                      "ms_per_call": round(dt * 1e3, 3), "Msps": round(S * T * ex.n_out / dt / 1e6, 1)}), flush=True)
    ex.close()


run(1, False)
for V in (2, 8, 64):
    run(V, False)
run(8, True) if S <= 512 else None
