python -m pytest tests/test_gpu_chain3.py tests/test_gpu_fuzz_chain.py tests/test_gpu_parity.py tests/test_gpu_params.py -q 2>&1 | grep -E "^(FAILED|E  +Assertion|E  +avdsp)|passed|failed" | cut -c1-300 | tail -12
python tools/float_tail_bench.py | tee gpurun_out/r2_float_tail_bench.jsonl
python bench.py --workload c3f --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3f', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
