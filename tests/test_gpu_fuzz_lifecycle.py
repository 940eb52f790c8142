"""One Executor through a random sequence of the C ABI's calls -- process (random lengths, so AUTO alternates between the kernels
of a program), process_range over two halves of the batch, reset with new seeds, a get_state / set_state round trip, a
reload_params with unchanged words -- mirrored on one oracle instance per stream.  Random programs of three generators, fixed
point and float formats 3 and 5 where the generator makes sense there.  Outputs of every call and the final state, bit for bit."""
import numpy as np
import pytest
import torch

from avdsp_b200 import Executor, synth
from oracle import wire
from test_gpu_parity import expected_state
from test_gpu_fuzz import random_program, nan_aware_equal
from test_gpu_fuzz_chain import random_chain_program
from test_gpu_fuzz_mix import random_mix_program

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(18))
def test_random_call_sequences(oracle_lib, seed):
    rng = np.random.default_rng(11000 + seed)
    fs = 48000
    which = seed % 3
    fmt = 2 if which == 0 else int(rng.choice([2, 3, 5] if which == 1 else [2, 3]))
    w = (random_program, random_chain_program, random_mix_program)[which](rng, fs, fmt)
    S = int(rng.choice([2, 6, 40]))
    seeds = np.arange(S, dtype=np.int32) * 3 + seed
    ex = Executor(w, fs, fmt, S, seeds=seeds, dither=24)
    gen = synth.pcm_float if fmt >= 5 else synth.pcm
    oracles = [oracle_lib.Oracle(w, fmt, fs, seed=int(seeds[s]), dither=24) for s in range(S)]
    log = []
    for step in range(int(rng.integers(4, 9))):
        op = str(rng.choice(["P", "P", "P", "Q", "R", "G", "L"]))
        log.append(op)
        if op in "PQ":
            T = int(rng.choice([1, 7, 64, 200, 1600 if S <= 6 else 300]))
            x = gen(str(rng.choice(["full", "noise", "impulse"])), S, T, max(ex.n_in, 1), fs)[:, :, : ex.n_in]
            if op == "P" or S < 2:
                y = ex.process(x)
            else:
                cut = int(rng.integers(1, S))
                xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
                ya = ex.process_range(xd[:cut].contiguous(), 0)
                yb = ex.process_range(xd[cut:].contiguous(), cut)
                torch.cuda.synchronize()
                y = np.concatenate([ya.cpu().numpy(), yb.cpu().numpy()], axis=0)
            log[-1] += f"{T}:{ex.last_kernel}"
            for s in range(S):
                ys = oracles[s].process(x[s])
                assert nan_aware_equal(y[s], ys, fmt >= 5), f"seed {seed} stream {s} after {log}: {np.count_nonzero(y[s] != ys)} samples differ\n" + "\n".join(wire.disassemble(w))
        elif op == "R":
            seeds = seeds + 100
            ex.reset(seeds=seeds, dither=24)
            oracles = [oracle_lib.Oracle(w, fmt, fs, seed=int(seeds[s]), dither=24) for s in range(S)]
        elif op == "G":
            s = int(rng.integers(0, S))
            ex.set_state(s, ex.get_state(s))
        else:
            ex.reload_params(w)
    for s in sorted({0, S - 1}):
        o = oracles[s]
        got, exp = ex.get_state(s), expected_state(ex, (o.data.copy(), o.aux(), o.code.copy()))
        assert nan_aware_equal(got, exp, fmt != 2), f"seed {seed} stream {s} after {log}: state differs at {np.nonzero(got != exp)[0][:8]}\n" + "\n".join(wire.disassemble(w))
