"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): batched kernel (32-, 64- and 128-thread variants,
cooperative merges), the lane-per-frame path and the large-frame path."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
c = synth.config(1, nframes=16); r = g.tessellate(c['xyz'], c['box'], 0.8, 3, want_tri=True, want_corr=True); assert (r['status'] == 0).all()
c = synth.config(3, nframes=6); r = g.tessellate(c['xyz'], c['box'], 0.8, 1, want_tri=True); assert (r['status'] == 0).all()
p = np.random.default_rng(0).random((2, 300, 2)); t, st = g.triangulate(p); assert (st == 0).all()
t, st, _ = g.triangulate_large(np.random.default_rng(1).random((3000, 2))); assert st == 0
print('sanitize case ok', os.environ.get('GTA_B200_LPF'))
