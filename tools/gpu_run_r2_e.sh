set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_params.py tests/test_gpu_shim.py -m gpu -q -x > $O/r2e_pytest.log 2>&1; tail -30 $O/r2e_pytest.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plugin_order or multi_device" >> $O/r2e_pytest.log 2>&1; tail -5 $O/r2e_pytest.log
bash tools/shim_latency.sh > $O/r2_shim_latency.jsonl 2> $O/r2_shim_latency.err; cat $O/r2_shim_latency.jsonl
