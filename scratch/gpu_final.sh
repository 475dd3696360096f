#!/bin/bash
# final single-GPU round: all GPU tests, bench (both arms), other configs, launch list, full capture of the star kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python scratch/measure_configs.py > gpurun_out/other_configs.json 2> gpurun_out/other_configs.err; tail -2 gpurun_out/other_configs.err
python scratch/star_time.py 100000 3 > gpurun_out/star_time.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scratch/star_time.py 100000 3 > gpurun_out/ncu_launch.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:star_kernel -s 1 -c 1 -o gpurun_out/star_full -f python scratch/star_time.py 16000 3 > gpurun_out/ncu_full.log 2>&1
python scratch/large_once.py > gpurun_out/large_once.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_large.csv env GTA_B200_PATH=2 python scratch/large_star.py 1000000 > gpurun_out/ncu_large.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_large.csv | tee gpurun_out/launches_large_summary.txt
python scratch/large_corr.py 1000 2>&1 | tee gpurun_out/large_corr.log
ls -la gpurun_out | head -40
