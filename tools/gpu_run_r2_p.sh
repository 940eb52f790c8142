python -m pytest tests/test_gpu_chain3.py tests/test_gpu_fuzz_chain.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_fuzz_mix.py -q 2>&1 | grep -E "^(FAILED|E  +Assertion|E  +avdsp)|passed|failed" | cut -c1-300 | tail -12
python bench.py --workload c3f --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3f', d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
python tools/kernel_survey.py 2>&1 | grep -E "f3|f4|f5|f6|dsptest" | cut -c1-110
