O=gpurun_out
python -m pytest tests/test_gpu_chain3.py -q -x 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain3 -s 3 -c 1 -f -o $O/r2_chain3_c3f python bench.py --workload c3f --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c3f.log 2>&1; tail -2 $O/ncu_c3f.log
