set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python benchmarks/ab_pipeline.py --only big --variants sync_b0,pipe_b0 > gpurun_out/r2e_ab.jsonl 2> gpurun_out/r2e_ab.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2e_sanitizer_$tool.log python __graft_entry__.py smoke > gpurun_out/r2e_sanitizer_$tool.out 2>&1; echo "$tool rc=$?" >> gpurun_out/r2e_sanitizer_$tool.out
done
tail -3 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_ab.jsonl; tail -5 gpurun_out/r2e_sanitizer_*.log
