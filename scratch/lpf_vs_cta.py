import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
cfg = int(sys.argv[1]); sizes = [int(x) for x in sys.argv[2].split(',')]
base = synth.config(cfg, nframes=2000)
N = base['xyz'].shape[1]
maxp = N + 4 + 4 * int(base['box'][:, :2].max() / 0.8)
for nf in sizes:
    reps = (nf + 1999) // 2000
    xyz = torch.from_numpy(np.tile(base['xyz'], (reps, 1, 1))[:nf]).cuda(); box = torch.from_numpy(np.tile(base['box'], (reps, 1))[:nf]).cuda()
    T = g.Tessellator(nf, N, maxp, 0.8, base['flags'])
    best = 1e9
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); T.run(xyz, box, 3); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print('cfg%d LPF=%s frames %d: %.1f ms -> %.0f frames/s' % (cfg, os.environ.get('GTA_B200_LPF'), nf, best * 1e3, nf / best), 'bad', int((T.status != 0).sum()))
