set -x; mkdir -p gpurun_out
O=gpurun_out
# launch list of the bench command itself (per-launch time + DRAM bytes), final code
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2y_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --parity-queries 8 > $O/r2y_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python benchmarks/summarize_ncu.py traffic $O/r2y_launches_bench.csv $O/r2y_step_traffic.json > $O/r2y_step_traffic.txt 2>&1
python benchmarks/summarize_ncu.py launches $O/r2y_launches_bench.csv > $O/r2y_launches_bench_summary.txt 2>&1
cat $O/r2y_launches_bench_summary.txt
# full capture of the largest int8 scan launch of the second search, final kernel
timeout 300 python benchmarks/profile_small_i8.py --queries 2514 --cta-group 2 > $O/r2y_plain_q2514.log 2>&1
CH=$(python -c "import json;print(json.loads(open('$O/r2y_plain_q2514.log').read().strip().splitlines()[-1])['chunks'])")
timeout 500 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s $((2*CH-1)) -c 1 -o $O/r2y_scan_q2514 -f python benchmarks/profile_small_i8.py --queries 2514 --cta-group 2 > $O/r2y_ncu_full_q2514.log 2>&1; echo "ncu full rc=$?"
ls -la $O/r2y*
