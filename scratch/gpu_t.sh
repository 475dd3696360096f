set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu --durations=5 2>&1 | tail -12 | tee gpurun_out/pytest_large.log
