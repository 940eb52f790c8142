"""What a silence tail costs in the float class (DSP_FORMAT 3, C3 at batch width): python tools/float_tail_bench.py
Streams whose cascades sit next to the float underflow threshold are re-executed by the interpreter (avdsp_dev.cuh fltGuard)."""
import sys, time, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from conftest import load_program
from avdsp_b200 import Executor, synth, KERNEL_GENERIC

w, fs = load_program("c3_peq16_f3_48k"), 48000
S, T = 65536, 4096
x = torch.from_numpy(synth.pcm("noise", S, T, 2, fs)).cuda()
zero = torch.zeros_like(x)
y = torch.empty((S, T, 2), dtype=torch.int32, device="cuda")


def timed(ex, inp, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): ex.process(inp, out=y)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3


ex = Executor(w, fs, 3, S, seeds=np.arange(S, dtype=np.int32))
ex.process(x, out=y); ex.process(x, out=y)
print(json.dumps({"case": "every stream plays noise", "kernel": ex.last_kernel, "ms_per_call": round(timed(ex, x), 3)}), flush=True)
# a fraction of the streams falls silent: their cascades decay (C3's peaking filters need ~25000 frames to reach 2^-64)
for frac in (0.03, 0.25, 1.0):
    ex = Executor(w, fs, 3, S, seeds=np.arange(S, dtype=np.int32))
    ex.process(x, out=y)
    mixed = x.clone()
    k = max(1, int(S * frac))
    idx = torch.arange(0, S, S // k, device="cuda")[:k]
    mixed[idx] = 0
    for _ in range(8): ex.process(mixed, out=y)           # 32768 frames of silence on those streams
    print(json.dumps({"case": f"{k} of {S} streams in a silence tail", "kernel": ex.last_kernel, "ms_per_call": round(timed(ex, mixed, 2), 3)}), flush=True)
ex = Executor(w, fs, 3, S, seeds=np.arange(S, dtype=np.int32))
ex.set_kernel(KERNEL_GENERIC)
ex.process(x, out=y)
print(json.dumps({"case": "interpreter, every stream plays noise", "kernel": ex.last_kernel, "ms_per_call": round(timed(ex, x, 2), 3)}), flush=True)
