set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 500 python benchmarks/ab_pipeline.py --only big --reps 6 --variants w768k_g35,w768k_g25,w768k_g50,w1536k_g35,w384k_g25,w768k_g18 > $O/r3a_ab.jsonl 2> $O/r3a_ab.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r3a_ab.jsonl'):
    d=json.loads(l)
    print('Q',d['Q'],'%-12s'%d['variant'],'total %.3f best %.3f scan %.3f chunks %d rescored %.2fM'%(d['total_ms'],d['best_ms'],d['scan_ms'],d['chunks'],d['rescored']/1e6))
PY
tail -3 $O/r3a_ab.err
