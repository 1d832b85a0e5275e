set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2i_tests.log 2>&1; echo "tests rc=$?" >> $O/r2i_tests.log
tail -4 $O/r2i_tests.log
timeout 400 python bench.py > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --config turn > $O/r2i_turn.json 2> $O/r2i_turn.err; echo "turn rc=$?"
timeout 300 python bench.py --config ksweep > $O/r2i_ksweep.json 2> $O/r2i_ksweep.err; echo "ksweep rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2i_ref.json 2> $O/r2i_ref.err; echo "ref rc=$?"
# launch list of one headline search with DRAM bytes (whole-step traffic)
timeout 300 python benchmarks/profile_small_i8.py --queries 2514 --cta-group 2 > $O/r2i_plain_q2514.log 2>&1
CH=$(python -c "import json;print(json.loads(open('$O/r2i_plain_q2514.log').read().strip().splitlines()[-1])['chunks'])")
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2i_launches_q2514.csv python benchmarks/profile_small_i8.py --queries 2514 --cta-group 2 > $O/r2i_ncu_launches.log 2>&1
python benchmarks/summarize_ncu.py traffic $O/r2i_launches_q2514.csv $O/r2i_step_traffic.json > $O/r2i_step_traffic.txt 2>&1
python benchmarks/summarize_ncu.py launches $O/r2i_launches_q2514.csv > $O/r2i_launches_q2514_summary.txt 2>&1
cat $O/r2i_step_traffic.txt
# full capture of the largest int8 scan launch of the second search
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s $((2*CH-1)) -c 1 -o $O/r2i_scan_q2514 -f python benchmarks/profile_small_i8.py --queries 2514 --cta-group 2 > $O/r2i_ncu_full_q2514.log 2>&1
timeout 300 python benchmarks/profile_small_i8.py --queries 1 --cta-group 1 > $O/r2i_plain_q1.log 2>&1
CH1=$(python -c "import json;print(json.loads(open('$O/r2i_plain_q1.log').read().strip().splitlines()[-1])['chunks'])")
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s $((2*CH1-1)) -c 1 -o $O/r2i_scan_q1 -f python benchmarks/profile_small_i8.py --queries 1 --cta-group 1 > $O/r2i_ncu_full_q1.log 2>&1
ls -la $O | tail -20
python - <<'PY'
import json
for f in ['r2i_bench','r2i_turn','r2i_ksweep','r2i_ref']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('kernel_ms_per_step'))
    except Exception as e:
        print(f,'FAILED',e)
PY
