set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
timeout 300 python scratch/star_time.py 50000 3 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_star.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab.log
