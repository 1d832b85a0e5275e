set -x; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warm or pipelined or int8_screen_path or exchange" > gpurun_out/r2g_tests_warm.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests_warm.log
tail -5 gpurun_out/r2g_tests_warm.log
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 4 --variants sync_b0,sync_w128k,sync_w256k,sync_w384k,sync_w768k,sync_w384k_g35,sync_w384k_g100,pipe_b0,pipe_w384k,pipe_w384k_s132x,pipe_x,pipe_s140x,pipe_s132x,pipe_s124x > gpurun_out/r2g_ab.jsonl 2> gpurun_out/r2g_ab.err
cat gpurun_out/r2g_ab.jsonl | cut -c1-330
tail -3 gpurun_out/r2g_ab.err
