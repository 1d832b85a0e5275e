set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "int8 or warm or golden or other_dimensions" > $O/r2k_tests.log 2>&1; echo "tests rc=$?" >> $O/r2k_tests.log
tail -4 $O/r2k_tests.log
timeout 600 python benchmarks/ab_pipeline.py --only big,small --reps 5 --variants v0,v1,v0_k1,v1_k1 > $O/r2k_ab.jsonl 2> $O/r2k_ab.err
cut -c1-300 $O/r2k_ab.jsonl
tail -3 $O/r2k_ab.err
