set -x; mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r3f_build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/r3f_tests.log 2>&1; echo "tests rc=$?" >> $O/r3f_tests.log
tail -5 $O/r3f_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r3f_smoke.log | cut -c1-200
timeout 500 python bench.py > $O/r3f_bench.json 2> $O/r3f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'], d['parity_check']['ok'])
PY
