#!/bin/bash
# usage: sweep.sh "0 4 16 32 64"  -> frames/s for each GTA_B200_COOP_NODES
for c in $1; do
  GTA_B200_COOP_NODES=$c timeout 300 python bench.py --steps 3 --warmup 3 --frames 20000 --no-cpu-baseline > gpurun_out/sw_$c.json 2> gpurun_out/sw_$c.err || tail -2 gpurun_out/sw_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/sw_$c.json')); print('coop_nodes=$c value %.0f e2e %.0f frames/s  frames_per_sm %s' % (d['value'], d['e2e']['value'], d['config']['frames_per_sm']))"
done
