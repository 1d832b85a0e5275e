set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "cta_pairs or auto_path or large_k or lazy or exchange or in_process or raw_c_abi" > $O/r2s_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s_tests.log
tail -15 $O/r2s_tests.log
timeout 400 python benchmarks/ab_pipeline.py --only big --reps 6 --variants def_tm,def_qs > $O/r2s_ab.jsonl 2> $O/r2s_ab.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2s_ab.jsonl'):
    d=json.loads(l)
    print('Q',d['Q'],'%-10s'%d['variant'],'total %.3f best %.3f scan %.3f chunks %d rescored %.2fM'%(d['total_ms'],d['best_ms'],d['scan_ms'],d['chunks'],d['rescored']/1e6))
PY
tail -5 $O/r2s_ab.err
