import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
nf = int(sys.argv[1]); reps = nf // 2000
c = synth.config(3, nframes=2000)
xyz = torch.from_numpy(np.tile(c['xyz'], (reps, 1, 1))).cuda(); box = torch.from_numpy(np.tile(c['box'], (reps, 1))).cuda()
T = g.Tessellator(nf, 1024, 1024 + 140, 0.8, 1)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); T.run(xyz, box, 3); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('LPF=%s frames %d: %.1f ms -> %.0f frames/s' % (os.environ.get('GTA_B200_LPF'), nf, (t1 - t0) * 1e3, nf / (t1 - t0)), 'bad status', int((T.status != 0).sum()))
want = g.tessellate(c['xyz'][:200], c['box'][:200], 0.8, 1)
print('max rel diff vs cta kernel', float(np.max(np.abs(T.area[:200].cpu().numpy() - want['area']) / want['area'])))
