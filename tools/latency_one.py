"""call latency of one stream x one period through Executor.process (device buffers): python tools/latency_one.py <program> <fmt> <fs> [frames]"""
import sys, time, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from conftest import load_program
from avdsp_b200 import Executor, synth
prog, fmt, fs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for T in ([int(sys.argv[4])] if len(sys.argv) > 4 else [64, 256, 1024]):
    ex = Executor(load_program(prog), fs, fmt, 1, seeds=[1])
    x = torch.from_numpy(synth.pcm("noise", 1, T, ex.n_in, fs)).cuda()
    y = torch.empty((1, T, ex.n_out), dtype=torch.int32, device="cuda")
    for _ in range(50): ex.process(x, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(400):
        t0 = time.perf_counter(); ex.process(x, out=y); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(json.dumps({"program": prog, "fmt": fmt, "frames": T, "kernel": ex.last_kernel, "p50_us": round(float(np.median(ts)), 1), "p99_us": round(float(np.percentile(ts, 99)), 1)}), flush=True)
