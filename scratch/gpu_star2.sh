set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_star.py -x -q > gpurun_out/pytest_star.log 2>&1; tail -5 gpurun_out/pytest_star.log
timeout 300 python scratch/star_time.py 100000 3 > gpurun_out/star_time.log 2>&1; cat gpurun_out/star_time.log
timeout 300 python scratch/star_time.py 100000 1 > gpurun_out/star_time1.log 2>&1; tail -1 gpurun_out/star_time1.log
bash scratch/gpu_star_ncu.sh 8000
