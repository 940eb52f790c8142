# Round-2 evidence on one B200 (gpurun): GPU tests, bench lines of every workload (both arms for the headline), the kernel
# survey, ncu launch lists and full captures of the dominant kernels.  Bench numbers come from the runs WITHOUT ncu.
set -x
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > $O/r2_pytest_gpu.log 2>&1; cat $O/r2_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_c2_reference.json 2> $O/r2_bench_c2_reference.err
python bench.py --steps 5 --warmup 3 > $O/r2_bench_c2_n1.json 2> $O/r2_bench_c2_n1.err
python bench.py --workload c5 --steps 5 --warmup 3 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
python bench.py --workload c3f --steps 5 --warmup 3 --no-e2e > $O/r2_bench_c3f.json 2> $O/r2_bench_c3f.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err
python bench.py --workload c4f --steps 3 --warmup 3 --no-e2e > $O/r2_bench_c4f.json 2> $O/r2_bench_c4f.err
python tools/kernel_survey.py > $O/r2_kernel_survey.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain2|k_chain3|k_int_peak|k_mix|k_fir|k_dag|k_generic|k_init' -c 40 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain3 -s 3 -c 1 -f -o $O/r2_chain3_c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_dag -s 2 -c 1 -f -o $O/r2_dag_dacfabriceo python tools/run_one.py ref_dacfabriceo 96000 > $O/ncu_dag_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mix_stream -s 3 -c 1 -f -o $O/r2_mix_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c5_full.log 2>&1
for f in $O/r2_bench_*.json; do echo $f; head -c 400 $f; echo; done
ls -la $O | tail -30
