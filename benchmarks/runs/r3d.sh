set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3d_tests.log
tail -15 gpurun_out/r3d_tests.log
