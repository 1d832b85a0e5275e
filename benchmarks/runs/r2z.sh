set -x; mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
timeout 300 python benchmarks/bench_inprocess.py --gpus $N > $O/r2z_inprocess_n$N.json 2> $O/r2z_inprocess_n$N.err; echo "inprocess rc=$?"
cat $O/r2z_inprocess_n$N.json; tail -3 $O/r2z_inprocess_n$N.err
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2z_n${N}_default.json 2> $O/r2z_n${N}_default.err; echo "rc=$?"
tail -3 $O/r2z_n${N}_default.err
N=$N python - <<'PY'
import json, os
n=os.environ['N']
d=json.loads(open('gpurun_out/r2z_n%s_default.json'%n).read().strip().splitlines()[-1])
s=d['stats']
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms_per_step'],3), 'pairs/rank', round(s['candidates_rescored_per_step']), 'chunks', s['n_chunks'], s['multi_gpu_phase_ms_max_over_ranks'], s['local_search_ms_per_rank'])
print({k:(round(v['ms_per_step'],2), v['path']) for k,v in d['secondary']['ksweep'].items()})
print({k:round(v['ms_per_search'],3) for k,v in d['secondary']['turn_latency'].items()})
print(d['parity_check']['ok'], d['clocks'])
PY
