set -x; mkdir -p gpurun_out
O=gpurun_out
N=4
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { # name, HAC_OPTIONS
  HAC_OPTIONS="$2" $TR bench.py --gpus $N --steps 8 --warmup 3 --no-secondary --no-cpu-baseline > $O/r2j_n${N}_$1.json 2> $O/r2j_n${N}_$1.err; echo "$1 rc=$?"
}
run default ""
run g100 "i8_chunk_growth_x100=100"
run g35 "i8_chunk_growth_x100=35"
run w0 "i8_warm_rows=0"
run w256k "i8_warm_rows=262144"
run w64k "i8_warm_rows=65536"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_n4_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d['stats']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms_per_step'],3), 'pairs/rank', round(s['candidates_rescored_per_step']), 'chunks', s['n_chunks'], s['multi_gpu_phase_ms_max_over_ranks'], s['local_search_ms_per_rank'])
    except Exception as e:
        print(f, 'FAILED', e)
PY
