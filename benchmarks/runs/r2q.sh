set -x; mkdir -p gpurun_out
O=gpurun_out
N=8
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2q_n8_default.json 2> $O/r2q_n8_default.err; echo "rc=$?"
timeout 300 python benchmarks/bench_inprocess.py --gpus $N > $O/r2q_inprocess_n$N.json 2> $O/r2q_inprocess_n$N.err; echo "inprocess rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_n8_default.json').read().strip().splitlines()[-1])
s=d['stats']
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms_per_step'],3), 'pairs/rank', round(s['candidates_rescored_per_step']), 'chunks', s['n_chunks'], s['multi_gpu_phase_ms_max_over_ranks'], s['local_search_ms_per_rank'])
print(open('gpurun_out/r2q_inprocess_n8.json').read())
PY
