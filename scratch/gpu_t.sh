set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_kept_api.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_large.log
