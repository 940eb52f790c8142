# latency of ONE stream x ONE period through the ALSA shim (the plugin's real operating point) and of the per-frame
# dspRuntime_2 compatibility call; JSON lines on stdout.  Run on the GPU box.
cd "$(dirname "$0")/.."
P=tests/golden/programs/c2_testrpi_xover_f2_192k.bin
head -c 65536 /dev/urandom > /tmp/shim_in.raw
for order in plugin canonical; do
  for period in 64 256 1024 4096; do
    shim/shim_harness $P 192000 s32 $period /tmp/shim_in.raw /tmp/shim_out.raw order=$order --latency 400 | grep period_frames | sed "s/^{/{\"order\": \"$order\", /"
  done
done
python - <<'PY'
import ctypes as C, json, time, sys
import numpy as np
sys.path.insert(0, ".")
from avdsp_b200 import _lib, program
L = _lib.lib()
w = program.load("tests/golden/programs/c2_testrpi_xover_f2_192k.bin")
buf = np.zeros(len(w) + 4096, np.int32); buf[:len(w)] = w
total = L.dspRuntimeInit(buf.ctypes.data, len(buf), 192000, 0, 31)
cores = []
for k in range(1, 9):
    p = L.dspFindCore(buf.ctypes.data, k)
    if not p: break
    cores.append(L.dspFindCoreBegin(p))
io = np.zeros(32, np.int32)
data = buf.ctypes.data + 4 * total
ts = []
for n in range(600):
    io[8], io[9] = 12345 * n, -777 * n
    t0 = time.perf_counter()
    for c in cores:
        L.dspRuntime_2(c, data, io.ctypes.data)
    ts.append((time.perf_counter() - t0) * 1e6)
ts = sorted(ts[100:])
print(json.dumps({"call": "dspRuntime_2 compatibility path, one frame = %d core calls" % len(cores), "p50_us": ts[len(ts)//2], "p99_us": ts[int(len(ts)*0.99)], "frames": len(ts)}))
PY
