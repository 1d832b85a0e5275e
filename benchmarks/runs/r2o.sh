set -x; mkdir -p gpurun_out
O=gpurun_out
N=8
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { # name, HAC_OPTIONS, extra args
  HAC_OPTIONS="$2" $TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline $3 > $O/r2o_n${N}_$1.json 2> $O/r2o_n${N}_$1.err; echo "$1 rc=$?"
}
run default "" ""
run g35 "i8_chunk_growth_x100=35" ""
run g100 "i8_chunk_growth_x100=100" ""
run w32k "i8_warm_rows=32768" ""
run w200k "i8_warm_rows=204800" ""
run qrecc "" "--config qrecc"
timeout 300 python benchmarks/bench_inprocess.py --gpus $N > $O/r2o_inprocess_n$N.json 2> $O/r2o_inprocess_n$N.err; echo "inprocess rc=$?"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2o_tests_multi.log 2>&1; echo "tests rc=$?" >> $O/r2o_tests_multi.log
tail -3 $O/r2o_tests_multi.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2o_n8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d['stats']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms_per_step'],3), 'pairs/rank', round(s['candidates_rescored_per_step']), 'chunks', s['n_chunks'], s['multi_gpu_phase_ms_max_over_ranks'], s['local_search_ms_per_rank'])
    except Exception as e:
        print(f, 'FAILED', e)
print(open('gpurun_out/r2o_inprocess_n8.json').read())
PY
tail -5 $O/r2o_inprocess_n8.err
