set -x
O=gpurun_out
mkdir -p $O
python tools/kernel_survey.py > $O/r2_kernel_survey.txt 2>&1; cat $O/r2_kernel_survey.txt
( time python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/r2g_pytest_gpu.log 2>&1; tail -20 $O/r2g_pytest_gpu.log
