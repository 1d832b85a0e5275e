set -x; mkdir -p gpurun_out
timeout 400 python benchmarks/bench_anisotropic.py > gpurun_out/r3b_aniso.json 2> gpurun_out/r3b_aniso.err; echo "rc=$?"
cat gpurun_out/r3b_aniso.json; tail -3 gpurun_out/r3b_aniso.err
timeout 300 python benchmarks/bench_anisotropic.py --k 1000 > gpurun_out/r3b_aniso_k1000.json 2> gpurun_out/r3b_aniso_k1000.err; echo "rc=$?"
cat gpurun_out/r3b_aniso_k1000.json; tail -3 gpurun_out/r3b_aniso_k1000.err
