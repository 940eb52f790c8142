# multi-GPU evidence (gpurun --gpus N): C2 weak scaling with the e2e arms, and BASELINE configs[2] as written: 65536 float streams
# (C3, DSP_FORMAT 3) SHARDED over the N GPUs (--scaling strong)
set -x
O=gpurun_out
N=${1:-2}
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > $O/r2_bench_c2_n$N.json 2> $O/r2_bench_c2_n$N.err
tail -2 $O/r2_bench_c2_n$N.err; head -c 600 $O/r2_bench_c2_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 --workload c3f --streams 65536 --scaling strong --no-e2e --no-cpu > $O/r2_bench_c3f_strong_n$N.json 2> $O/r2_bench_c3f_strong_n$N.err
tail -2 $O/r2_bench_c3f_strong_n$N.err; head -c 600 $O/r2_bench_c3f_strong_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 --workload c3 --streams 65536 --scaling strong --no-e2e --no-cpu > $O/r2_bench_c3_strong_n$N.json 2> $O/r2_bench_c3_strong_n$N.err
head -c 300 $O/r2_bench_c3_strong_n$N.json; echo
