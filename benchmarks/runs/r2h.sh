set -x; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -4 gpurun_out/r2h_tests.log
TR="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2h_bench_n2_sync.json 2> gpurun_out/r2h_bench_n2_sync.err; echo "sync rc=$?"
HAC_I8_PIPELINE=1 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2h_bench_n2_pipe.json 2> gpurun_out/r2h_bench_n2_pipe.err; echo "pipe rc=$?"
for f in gpurun_out/r2h_bench_n2_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'], d['stats'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
tail -3 gpurun_out/r2h_bench_n2_pipe.err
