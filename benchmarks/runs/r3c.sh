set -x; mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --config qrecc --steps 5 --warmup 3 --no-cpu-baseline > $O/r3c_qrecc_n${N}.json 2> $O/r3c_qrecc_n${N}.err; echo "rc=$?"
tail -3 $O/r3c_qrecc_n${N}.err
N=$N python - <<'PY'
import json, os
n=os.environ['N']
d=json.loads(open('gpurun_out/r3c_qrecc_n%s.json'%n).read().strip().splitlines()[-1])
s=d['stats']
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms_per_step'],3), d['roofline']['frac'], 'pairs/rank', round(s['candidates_rescored_per_step']), 'chunks', s['n_chunks'], s['multi_gpu_phase_ms_max_over_ranks'], s['hbm_fp32_gb'], s['hbm_int8_gb'], s['hbm_f16_gb'])
print(d['parity_check']['ok'], d['clocks'])
PY
