import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
from conftest import load_program
from avdsp_b200 import Executor
ex = Executor(load_program("ref_dacdiy1"), 192000, 2, 4096)
print("\n".join(l for l in ex.trace.splitlines() if "geometry" in l or "kernel" in l))
