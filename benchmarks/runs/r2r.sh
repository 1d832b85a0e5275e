set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r2r_tests.log 2>&1; echo "tests rc=$?" >> $O/r2r_tests.log
tail -15 $O/r2r_tests.log
timeout 500 python benchmarks/ab_k1000.py --ks 1000,250 --reps 2 > $O/r2r_k1000.jsonl 2> $O/r2r_k1000.err; echo "k1000 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2r_k1000.jsonl'):
    d=json.loads(l)
    if 'failed' in d: print(d); continue
    print('k',d['k'],'%-16s'%d['variant'],'total %.2f best %.2f scan %.2f chunks %d launches %d rescored %.1fM path %d same %s'%(d['total_ms'],d['best_ms'],d['scan_ms'],d['chunks'],d['launches'],d['rescored']/1e6,d['path'],d['identical_results']))
PY
tail -5 $O/r2r_k1000.err
