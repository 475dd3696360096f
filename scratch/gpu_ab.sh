set -x
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
for rep in 1 2; do
for lib in g_tessla_b200/libgtessla_b200.so scratch/libdbg_old.so; do
  GTA_B200_LIB=$PWD/$lib timeout 300 python scratch/star_time.py 50000 3 2>&1 | tail -1 | cut -c1-100 | sed "s|^|$lib |" | tee -a gpurun_out/ab.log
done; done
