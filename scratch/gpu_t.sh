set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kept_api.py -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_large.log
