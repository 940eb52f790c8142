# BASELINE configs[2] as written: 65536 float streams (C3, DSP_FORMAT 3) SHARDED over the N GPUs (--scaling strong); fixed point beside it
O=gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --workload c3f --streams 65536 --scaling strong --no-e2e --no-cpu > $O/r2_bench_c3f_strong_n$N.json 2> $O/r2_bench_c3f_strong_n$N.err
tail -1 $O/r2_bench_c3f_strong_n$N.err; head -c 400 $O/r2_bench_c3f_strong_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 --workload c3 --streams 65536 --scaling strong --no-e2e --no-cpu > $O/r2_bench_c3_strong_n$N.json 2> $O/r2_bench_c3_strong_n$N.err
head -c 400 $O/r2_bench_c3_strong_n$N.json; echo
