set -x
O=gpurun_out
python -m pytest tests/test_gpu_chain3.py -q -x 2>&1 | tail -30
python -m pytest tests/test_gpu_parity.py -q -x -k "c3_peq16 or full_width or c2_" 2>&1 | tail -3
for w in c2 c3 c3f; do
python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu > $O/r2k_$w.json 2> $O/r2k_$w.err; tail -1 $O/r2k_$w.err; python -c "
import json; d=json.load(open('$O/r2k_$w.json')); print('$w', d['ms_per_step'], d['run']['kernel'], d['roofline']['frac'])"
done
