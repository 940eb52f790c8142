set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/r1_bench_c4.json 2> gpurun_out/r1_bench_c4.err; tail -c 600 gpurun_out/r1_bench_c4.json
python bench.py --workload c4f --steps 3 --warmup 3 > gpurun_out/r1_bench_c4f.json 2> gpurun_out/r1_bench_c4f.err; tail -c 300 gpurun_out/r1_bench_c4f.json
python bench.py --workload c4f --kernel fir_tc --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_bench_c4f_tc.json 2> gpurun_out/r1_bench_c4f_tc.err
python bench.py --workload c4 --kernel fir --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1_bench_c4_scalar.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_firtc -s 3 -c 1 -f -o gpurun_out/r1_firtc_i8_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_firtc -s 3 -c 1 -f -o gpurun_out/r1_firtc_tf32_c4f python bench.py --workload c4f --kernel fir_tc --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4f_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fir -s 6 -c 1 -f -o gpurun_out/r1_fir_f32_c4f python bench.py --workload c4f --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4f_exact.log 2>&1
ls -la gpurun_out | tail -12
