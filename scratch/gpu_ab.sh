set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for t in 64 128 256; do
  GTA_B200_STAR_THREADS=$t timeout 300 python scratch/star_time.py 100000 1 2>&1 | tail -1 | cut -c1-110 | sed "s/^/T=$t cfg1 /" | tee -a gpurun_out/ab.log
done
GTA_B200_PATH=1 timeout 300 python scratch/star_time.py 100000 1 2>&1 | tail -1 | cut -c1-110 | sed "s/^/DC cfg1 /" | tee -a gpurun_out/ab.log
