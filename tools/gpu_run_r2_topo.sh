# host topology of the box + multi-GPU bench (run with gpurun --gpus N)
set -x
O=gpurun_out
N=${1:-2}
mkdir -p $O
( nproc; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
  nvidia-smi topo -m
  for b in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b2=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo $b numa=$(cat /sys/bus/pci/devices/$b2/numa_node 2>/dev/null); done
  python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"
  grep -i "cpus_allowed_list\|mems_allowed_list" /proc/self/status ) > $O/r2_topology_n$N.txt 2>&1
cat $O/r2_topology_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $O/r2_bench_c2_n$N.json 2> $O/r2_bench_c2_n$N.err
tail -5 $O/r2_bench_c2_n$N.err; cat $O/r2_bench_c2_n$N.json
