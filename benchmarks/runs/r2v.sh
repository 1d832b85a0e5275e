set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2v_tests.log 2>&1; echo "tests rc=$?" >> $O/r2v_tests.log
tail -6 $O/r2v_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2v_smoke.log 2>&1; tail -2 $O/r2v_smoke.log
timeout 500 python bench.py > $O/r2v_bench.json 2> $O/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['clocks'])
print({k:round(v['ms_per_search'],3) for k,v in d['secondary']['turn_latency'].items()})
print({k:(round(v['ms_per_step'],2), v['path']) for k,v in d['secondary']['ksweep'].items()})
print(d['cpu_baseline'])
PY
