# A/B of star-kernel builds: GPU star tests on the in-tree library first, then star_time on every library named in LIBS
set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
timeout 600 python -m pytest ${TESTS:-tests/test_gpu_star.py tests/test_gpu_parity.py} -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_star.log
for rep in 1 2; do
for lib in $LIBS; do
  GTA_B200_LIB=$PWD/$lib timeout 300 python scratch/star_time.py 50000 3 2>&1 | tail -1 | cut -c1-110 | sed "s|^|$lib |" | tee -a gpurun_out/ab.log
done; done
