O=gpurun_out
( time python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $O/r2m_pytest_gpu.log 2>&1; cat $O/r2m_pytest_gpu.log
