set -x
mkdir -p gpurun_out
for t in 256 320 384; do
  GTA_B200_STAR_THREADS=$t timeout 300 python scratch/star_time.py 50000 3 2>&1 | tail -1 | sed "s/^/T=$t /" | tee -a gpurun_out/star_threads.log
done
