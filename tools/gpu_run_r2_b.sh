# round 2, call B: GPU test suite after the NaN / multi-device / stream-order changes, a short bench line, compute-sanitizer
set -x
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -30 ) > $O/r2b_pytest_gpu.log 2>&1; tail -35 $O/r2b_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err; tail -3 $O/r2b_bench_c2.err; cat $O/r2b_bench_c2.json
# race / barrier checks on the hand-synchronised kernels (k_chain3: spin-wait on per-part tile counters, split bar.arrive / bar.sync)
S=/usr/local/cuda/bin/compute-sanitizer
( timeout 900 $S --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_chain3.py -q -x -k "ragged and noise or planar and 96" 2>&1 | tail -25 ) > $O/r2_sanitizer_racecheck_chain3.log 2>&1; tail -8 $O/r2_sanitizer_racecheck_chain3.log
( timeout 900 $S --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_chain3.py -q -x -k "ragged and noise or planar and 96" 2>&1 | tail -25 ) > $O/r2_sanitizer_synccheck_chain3.log 2>&1; tail -8 $O/r2_sanitizer_synccheck_chain3.log
( timeout 900 $S --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "test_golden_vectors_auto_kernel and (c5_noise or c2_noise or c4_f3_noise or dacdiy1_48k or allops_gen_f2_48k)" 2>&1 | tail -25 ) > $O/r2_sanitizer_memcheck_kernels.log 2>&1; tail -8 $O/r2_sanitizer_memcheck_kernels.log
( timeout 900 $S --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "test_golden_vectors_auto_kernel and (c5_noise or c4_f3_noise or dacdiy1_48k)" 2>&1 | tail -25 ) > $O/r2_sanitizer_racecheck_mix_fir_chain2.log 2>&1; tail -8 $O/r2_sanitizer_racecheck_mix_fir_chain2.log
