"""cfg 3 (or CFG), N device-resident frames through gta_b200_tessellate_device: star path vs divide and conquer.  python scratch/star_time.py N [cfg]"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
nf = int(sys.argv[1]); cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
base = 2000
c = synth.config(cfg, nframes=base)
reps = max(1, nf // base); nf = reps * base
xyz = torch.from_numpy(np.tile(c['xyz'], (reps, 1, 1))).cuda(); box = torch.from_numpy(np.tile(c['box'], (reps, 1))).cuda()
n = c['xyz'].shape[1]
maxp = n + 4 + 4 * int(c['box'][:, :2].max() / c['espace'] + 1)
T = g.Tessellator(nf, n, maxp, c['espace'], c['flags'])
for mode in (g.PATH_AUTO, g.PATH_DC) if '--dc' in sys.argv else (g.PATH_AUTO,):
    g.set_path(mode)
    for it in range(3):
        g.stage_ms(reset=True); g.star_counters(reset=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); T.run(xyz, box, 3); torch.cuda.synchronize(); t1 = time.perf_counter()
        print('path %d frames %d: %.2f ms -> %.0f frames/s' % (mode, nf, (t1 - t0) * 1e3, nf / (t1 - t0)), 'bad status', int((T.status != 0).sum()),
              'stages', {k: round(v, 3) for k, v in g.stage_ms().items()}, 'certified/fallback', g.star_counters())
    if mode == g.PATH_AUTO: a_star = T.area.cpu().numpy().copy()
    else: print('max rel diff star vs dc', float(np.max(np.abs(T.area.cpu().numpy() - a_star) / a_star)))
