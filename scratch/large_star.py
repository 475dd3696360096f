"""One large frame through the star engine (GTA_B200_PATH=2), sizes on the command line; GTA_B200_TRACE=1 prints the kernel time."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
g.set_path(g.PATH_STAR)
for n in [int(a) for a in sys.argv[1:]] or [3, 100, 100000, 1000000]:
    p = synth.uniform_patch(n, seed=4) if n >= 1000 else np.random.default_rng(4).random((n, 2)) * 50
    for it in range(2):
        t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
    print('n %d: %d triangles, status %d, engine %d, device %.1f ms, with copies %.1f ms' % (n, len(tri), st, g.lib().gta_b200_large_last_engine(), ms, (t1 - t0) * 1e3), flush=True)
