"""One large frame through the star engine (GTA_B200_PATH=2), sizes on the command line; GTA_B200_TRACE=1 prints the kernel time."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
g.set_path(g.PATH_STAR)
for n in [int(a) for a in sys.argv[1:] if a.isdigit()] or [3, 100, 100000, 1000000]:
    p = synth.uniform_patch(n, seed=4) if n >= 1000 else np.random.default_rng(4).random((n, 2)) * 50
    if '--box' in sys.argv:        # points along the bounding box, one per cell side (what -corr adds to a frame): the hull is the box
        lo, hi = p.min(0), p.max(0); m = int(np.sqrt(n)); tx = np.linspace(lo[0], hi[0], m); ty = np.linspace(lo[1], hi[1], m)[1:-1]
        p = np.concatenate([p, np.stack([tx, np.full(m, lo[1])], 1), np.stack([tx, np.full(m, hi[1])], 1), np.stack([np.full(m - 2, lo[0]), ty], 1), np.stack([np.full(m - 2, hi[0]), ty], 1)])
    for it in range(2):
        t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
    print('n %d: %d triangles, status %d, engine %d, device %.1f ms, with copies %.1f ms' % (n, len(tri), st, g.lib().gta_b200_large_last_engine(), ms, (t1 - t0) * 1e3), flush=True)
