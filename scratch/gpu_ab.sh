set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
GTA_B200_TRACE=1 timeout 500 python scratch/large_star.py 100000 1000000 2>&1 | grep -v chunk | tee -a gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab.log
