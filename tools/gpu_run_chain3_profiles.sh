# Evidence for the headline kernel (k_chain3 on C2) on one B200 (gpurun): bench line with e2e + cpu baseline, the A/B line of
# k_chain2, ncu launch list and one full capture.  Bench numbers come from the runs WITHOUT ncu.
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/r1_bench_c2_n1.json 2> $O/r1_bench_c2_n1.err
python bench.py --steps 5 --warmup 3 --kernel chain_v2 --no-e2e --no-cpu > $O/r1_bench_c2_chain2.json 2> $O/r1_bench_c2_chain2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_chain2|k_chain3|k_int_peak|k_mix|k_fir' -c 40 --csv --log-file $O/r1_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain3 -s 3 -c 1 -f -o $O/r1_chain3_c2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_c2_full.log 2>&1
ls -la $O | tail -6
