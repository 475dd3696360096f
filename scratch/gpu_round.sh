#!/bin/bash
# one GPU call: tests, bench (both arms), launch list, full capture of the star kernel (each ncu pass only after its command ran clean)
set -x
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log; fi
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python scratch/star_time.py 100000 3 > gpurun_out/star_time.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scratch/star_time.py 100000 3 > gpurun_out/ncu_launch.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:star_kernel -s 1 -c 1 -o gpurun_out/star_full -f python scratch/star_time.py 16000 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
