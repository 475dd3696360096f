import sys, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
c = synth.config(1, nframes=500)
reps = 240
xyz = np.tile(c["xyz"], (reps, 1, 1)); box = np.tile(c["box"], (reps, 1))
n = c["xyz"].shape[1]
maxp = n + 4 + 4 * int(c["box"][:, :2].max() / c["espace"] + 1)
T = g.Tessellator(len(xyz), n, maxp, c["espace"], c["flags"])
g.star_counters(reset=True)
T.run(torch.from_numpy(xyz).cuda(), torch.from_numpy(box).cuda(), 3); torch.cuda.synchronize()
print('device: certified/fallback', g.star_counters(reset=True))
area = T.area.cpu().numpy().reshape(reps, -1)
print('device self-consistent', bool((area == area[0]).all()))
for k in range(2):
    r = g.tessellate(xyz, box, c["espace"], c["flags"])
    print('host: certified/fallback', g.star_counters(reset=True), 'launches', r["launches"])
    a = r["area"].reshape(reps, -1)
    d = a != area
    print('host differs in', int(d.sum()), 'entries; replicas affected', np.unique(np.nonzero(d)[0])[:20], 'frames', np.unique(np.nonzero(d)[1])[:20])
    print('host self-consistent', bool((a == a[0]).all()))
