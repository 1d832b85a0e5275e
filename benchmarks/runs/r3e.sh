set -x; mkdir -p gpurun_out
O=gpurun_out
N=8
timeout 200 python benchmarks/bench_inprocess.py --gpus $N --pinned > $O/r3e_inprocess_n${N}_pinned.json 2> $O/r3e_inprocess_n${N}_pinned.err; echo "rc=$?"
cat $O/r3e_inprocess_n${N}_pinned.json; tail -3 $O/r3e_inprocess_n${N}_pinned.err
