set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 6 --variants w768k_g35,w768k_g25,w768k_g18,w768k_g50,w1536k_g35,w1536k_g25,w384k_g25,w3072k_g25,pipe_w768k,sync_w0 > $O/r2l_ab.jsonl 2> $O/r2l_ab.err
cut -c1-330 $O/r2l_ab.jsonl
tail -3 $O/r2l_ab.err
