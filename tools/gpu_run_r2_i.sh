set -x
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dag -s 2 -c 1 -f -o $O/r2_dag_lv6 python tools/run_one.py ref_crossoverLV6 96000 > $O/ncu_dag_lv6.log 2>&1; tail -2 $O/ncu_dag_lv6.log
