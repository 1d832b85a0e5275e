set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "int8 or warm or exchange or in_process or golden" > $O/r2p_tests.log 2>&1; echo "tests rc=$?" >> $O/r2p_tests.log
tail -3 $O/r2p_tests.log
timeout 300 python benchmarks/ab_pipeline.py --only big,small --reps 6 --variants sync_b0,w768k_g35 > $O/r2p_ab.jsonl 2> $O/r2p_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/r2p_ab.jsonl'):
    d=json.loads(l)
    print("Q %-5d %-12s total %.3f best %.3f scan %.3f tail %.3f chunks %d launches %d" % (d['Q'], d['variant'], d['total_ms'], d['best_ms'], d['scan_ms'], d['tail_ms'], d['chunks'], d['launches']))
PY
