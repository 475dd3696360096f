set -x
mkdir -p gpurun_out
python scratch/star_time.py ${1:-8000} 3 > gpurun_out/star_time_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:star_kernel -s 1 -c 1 -o gpurun_out/star_full -f python scratch/star_time.py ${1:-8000} 3 > gpurun_out/ncu_star.log 2>&1
tail -3 gpurun_out/ncu_star.log
ls -la gpurun_out/*.ncu-rep
