set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_shim.py -m gpu -q -x > $O/r2d_pytest_shim.log 2>&1; tail -30 $O/r2d_pytest_shim.log
bash tools/shim_latency.sh > $O/r2_shim_latency.jsonl 2> $O/r2_shim_latency.err; cat $O/r2_shim_latency.jsonl; tail -3 $O/r2_shim_latency.err
