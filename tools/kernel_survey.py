import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from conftest import CASES, load_program
from avdsp_b200 import Executor, synth
# which kernel AUTO picks for every test program at batch width (4096 streams), and what it sustains (device-resident PCM)
seen = set()
for prog, fmt, fs in CASES:
    if (prog, fmt) in seen or "allops" in prog:
        continue
    seen.add((prog, fmt))
    w = load_program(prog)
    S, T = 4096, 4800
    ex = Executor(w, fs, fmt, S)
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    x = torch.from_numpy(gen("noise", S, T, ex.n_in, fs)).cuda()
    y = torch.empty((S, T, ex.n_out), dtype=torch.int32, device="cuda")
    ex.process(x, out=y); torch.cuda.synchronize()
    t0 = time.perf_counter(); ex.process(x, out=y); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    why = [l for l in ex.trace.splitlines() if "not used" in l or "chain kernel:" in l]
    print(f"{prog:32s} fmt{fmt} fs={fs:6d} kernel={ex.last_kernel:8s} {S*T*ex.n_out/dt/1e9:8.2f} G ch-samples/s  ({dt*1e3:7.1f} ms) in={ex.n_in} out={ex.n_out} {why[-1][:90] if why else ''}")
