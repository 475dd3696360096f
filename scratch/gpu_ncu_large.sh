set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ls_star_kernel -s 1 -c 1 -o gpurun_out/ls_full -f python scratch/large_star.py 100000 > gpurun_out/ncu_ls.log 2>&1
tail -3 gpurun_out/ncu_ls.log
