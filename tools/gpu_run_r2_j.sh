set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_chain3.py -q -x 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py -q -x -k "c3_peq16_f2 or full_width" 2>&1 | tail -3
for k in auto chain_v2 chain_v3; do
python bench.py --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu --kernel $k > $O/r2j_c3_$k.json 2> $O/r2j_c3_$k.err; tail -1 $O/r2j_c3_$k.err; python -c "
import json; d=json.load(open('$O/r2j_c3_$k.json')); print('C3 $k', d['ms_per_step'], d['run']['kernel'], d['roofline']['frac'])"
done
