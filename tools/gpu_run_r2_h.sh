set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dag.py -m gpu -q -x > $O/r2h_pytest_dag.log 2>&1; tail -5 $O/r2h_pytest_dag.log
python tools/kernel_survey.py 2>&1 | grep -E "dag|dacdiy|lr2" > $O/r2h_survey.txt; cat $O/r2h_survey.txt
