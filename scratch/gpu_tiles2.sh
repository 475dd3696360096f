set -x
mkdir -p gpurun_out; rm -f gpurun_out/large_tiles.log
timeout 600 python -m pytest tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_large.log
GTA_B200_TRACE=1 timeout 300 python scratch/large_star.py 100000 1000000 4000000 2>&1 | grep -v "chunk cycles" | sed "s/^/tiles: /" | tee -a gpurun_out/large_tiles.log
GTA_B200_LARGE_TILES=0 GTA_B200_TRACE=1 timeout 300 python scratch/large_star.py 1000000 2>&1 | grep -v "chunk cycles" | sed "s/^/no tiles: /" | tee -a gpurun_out/large_tiles.log
GTA_B200_DEFER_PER=4 GTA_B200_TRACE=1 timeout 300 python scratch/large_star.py 1000000 2>&1 | grep -v "chunk cycles" | sed "s/^/per 4: /" | tee -a gpurun_out/large_tiles.log
