set -x; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 12 --variants v1,v2 > $O/r2m_ab12.jsonl 2> $O/r2m_ab.err
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 12 --variants v2,v1 > $O/r2m_ab21.jsonl 2>> $O/r2m_ab.err
timeout 600 python benchmarks/ab_pipeline.py --only big --reps 8 --variants v0,v1,v2 > $O/r2m_ab012.jsonl 2>> $O/r2m_ab.err
tail -3 $O/r2m_ab.err
