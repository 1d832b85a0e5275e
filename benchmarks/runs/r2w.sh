set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --config ksweep --steps 5 --warmup 3 > $O/r2w_ksweep.json 2> $O/r2w_ksweep.err; echo "ksweep rc=$?"
timeout 300 python bench.py --config turn --steps 20 --warmup 3 > $O/r2w_turn.json 2> $O/r2w_turn.err; echo "turn rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2w_ref.json 2> $O/r2w_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_ksweep.json').read().strip().splitlines()[-1])
print({k:(round(v['ms_per_step'],2), v['path'], round(v['rescored_pairs']/1e6,2)) for k,v in d['sweep'].items()}, d['parity_check']['ok'], d['stats'])
d=json.loads(open('gpurun_out/r2w_turn.json').read().strip().splitlines()[-1])
print({k:(round(v['ms_per_search'],3), round(v['scan_ms'],3), round(v['hbm_gbs_streamed'])) for k,v in d['batches'].items()}, d['roofline']['frac'], d['clocks'])
d=json.loads(open('gpurun_out/r2w_ref.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['wall_s'])
PY
