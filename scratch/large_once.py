"""cfg 4 once (1M points in one frame) -- the command the ncu launch list of the large path is taken on."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
p = synth.uniform_patch(1_000_000, seed=4)
g.triangulate_large(p[:100000])
for it in range(3):
    t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
    print('cfg4: %d triangles, status %d, device %.1f ms, with copies %.1f ms' % (len(tri), st, ms, (t1 - t0) * 1e3))
