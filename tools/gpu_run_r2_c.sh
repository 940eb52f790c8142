set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device or ranges_on_two or chain3_full_width or allops_gen_f3 or allops_gen_f5" > $O/r2c_pytest.log 2>&1; head -80 $O/r2c_pytest.log; tail -5 $O/r2c_pytest.log
