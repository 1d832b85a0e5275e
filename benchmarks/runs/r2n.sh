set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2n_tests.log 2>&1; echo "tests rc=$?" >> $O/r2n_tests.log
tail -4 $O/r2n_tests.log
timeout 400 python bench.py > $O/r2n_bench.json 2> $O/r2n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['clocks'])
print(d['stats'])
print({k:v['ms_per_search'] for k,v in d['secondary']['turn_latency'].items()})
print({k:(v['ms_per_step']) for k,v in d['secondary']['ksweep'].items()})
PY
