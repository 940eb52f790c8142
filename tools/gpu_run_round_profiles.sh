# Round-end evidence run on one B200 (gpurun): GPU tests, bench lines of every workload, ncu launch lists + full captures.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/r1_bench_c2_n1.json 2> $O/r1_bench_c2_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r1_bench_c2_reference.json 2> $O/r1_bench_c2_reference.err
python bench.py --workload c5 --steps 5 --warmup 3 > $O/r1_bench_c5.json 2> $O/r1_bench_c5.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e > $O/r1_bench_c3.json 2> $O/r1_bench_c3.err
python bench.py --workload c3f --steps 5 --warmup 3 --no-e2e > $O/r1_bench_c3f.json 2> $O/r1_bench_c3f.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain2|k_chain3|k_int_peak|k_mix|k_fir' -c 40 --csv --log-file $O/r1_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain3 -s 3 -c 1 -f -o $O/r1_chain3_c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain2|k_chain3|k_int_peak|k_mix|k_fir' -c 40 --csv --log-file $O/r1_launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mix_stream -s 3 -c 1 -f -o $O/r1_mix_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c5_full.log 2>&1
ls -la $O | tail -14
