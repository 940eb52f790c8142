"""one program at batch width, a few launches (for ncu captures): python tools/run_one.py <program> <fs> [streams] [frames] [kernel]"""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from conftest import load_program
import avdsp_b200
from avdsp_b200 import Executor, synth
prog, fs = sys.argv[1], int(sys.argv[2])
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
T = int(sys.argv[4]) if len(sys.argv) > 4 else 4800
w = load_program(prog)
fmt = 2 if (int(w[6]) & 0xFFFF) else 3
ex = Executor(w, fs, fmt, S, seeds=np.arange(S, dtype=np.int32))
if len(sys.argv) > 5:
    ex.set_kernel(getattr(avdsp_b200, "KERNEL_" + sys.argv[5].upper()))
x = torch.from_numpy(synth.pcm("noise", S, T, ex.n_in, fs)).cuda()
y = torch.empty((S, T, ex.n_out), dtype=torch.int32, device="cuda")
for _ in range(3):
    ex.process(x, out=y)
torch.cuda.synchronize()
t0 = time.perf_counter(); ex.process(x, out=y); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{prog} kernel={ex.last_kernel}{ex.last_chain_variant or ''} {S}x{T}: {dt*1e3:.3f} ms, {S*T*ex.n_out/dt/1e9:.2f} G ch-samples/s")
