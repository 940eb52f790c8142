O=gpurun_out
python -m pytest tests/test_gpu_chain3.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -x 2>&1 | tail -5
python bench.py --workload c3f --steps 10 --warmup 3 2>&1 | tail -1 > $O/r2_bench_c3f.json; cut -c1-200 $O/r2_bench_c3f.json
./tools/microbench_f32x2 > $O/r2_microbench_f32x2.jsonl; grep -E '"warps_per_subpartition": (1|2),' $O/r2_microbench_f32x2.jsonl | cut -c1-200
