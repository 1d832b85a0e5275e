set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_pairs or warm" > $O/r2l_tests.log 2>&1; echo "tests rc=$?" >> $O/r2l_tests.log
tail -3 $O/r2l_tests.log
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 6 --variants v0,v1,w768k_g35,w768k_g25,w768k_g18,w768k_g50,w1536k_g35,w1536k_g25,w384k_g25,w3072k_g25,pipe_w768k,sync_w0 > $O/r2l_ab.jsonl 2> $O/r2l_ab.err
tail -3 $O/r2l_ab.err
