set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dag.py -m gpu -q -x > $O/r2f_pytest_dag.log 2>&1; tail -40 $O/r2f_pytest_dag.log
