O=gpurun_out
python -m pytest tests/test_gpu_params.py -q -x 2>&1 | tail -3
python tools/variants_bench.py 4096 4800 | tee $O/r2_variants_bench.jsonl
python tools/variants_bench.py 512 1024 | tee -a $O/r2_variants_bench.jsonl
