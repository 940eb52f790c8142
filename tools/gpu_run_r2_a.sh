# round 2, call A: GPU test suite with the new all-opcode goldens + the DFMA co-issue microbenchmark
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/r2a_pytest_gpu.log 2>&1; tail -45 $O/r2a_pytest_gpu.log
cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o microbench_dfma microbench_dfma.cu 2>/dev/null; ./microbench_dfma > ../$O/r2_microbench_dfma.jsonl; cd ..
cat $O/r2_microbench_dfma.jsonl
