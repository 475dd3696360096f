import os, sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200.api as api
api.lib_path = lambda: os.path.abspath(sys.argv[1])
from g_tessla_b200 import synth
c = synth.config(3, nframes=4000)
xyz = np.tile(c['xyz'], (5, 1, 1)); box = np.tile(c['box'], (5, 1))
api.tessellate(xyz[:4000], box[:4000], 0.8, 1)
r = api.tessellate(xyz, box, 0.8, 1)
print(sys.argv[1], 'kernel frames/s %.0f' % (len(xyz) / (r['kernel_ms'] / 1e3)), 'status ok', bool((r['status'] == 0).all()))
