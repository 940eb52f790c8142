set -x
O=gpurun_out
mkdir -p $O
python tools/run_one.py ref_dacfabriceo 96000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dag -s 2 -c 1 -f -o $O/r2_dag_fab python tools/run_one.py ref_dacfabriceo 96000 > $O/ncu_dag_fab.log 2>&1; tail -3 $O/ncu_dag_fab.log
