import sys, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
c = synth.config(1, nframes=500)
n = c["xyz"].shape[1]
xyz = torch.from_numpy(c["xyz"]).cuda(); box = torch.from_numpy(c["box"]).cuda()
outs = {}
for maxp in (100, 104, 140):
    T = g.Tessellator(500, n, maxp, c["espace"], c["flags"])
    T.run(xyz, box, 3); torch.cuda.synchronize()
    outs[maxp] = T.area.cpu().numpy().copy()
    T.run(xyz, box, 3); torch.cuda.synchronize()
    print(maxp, 'repeat equal', np.array_equal(outs[maxp], T.area.cpu().numpy()), 'vs 100 equal', np.array_equal(outs[maxp], outs[100]), 'ndiff', int((outs[maxp] != outs[100]).sum()))
r = g.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
for maxp in outs: print('host vs dev', maxp, int((r["area"] != outs[maxp]).sum()), float(np.max(np.abs(r["area"] - outs[maxp]) / outs[maxp])))
r2 = g.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
print('host repeat equal', np.array_equal(r["area"], r2["area"]))
g.set_path(g.PATH_DC)
r3 = g.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
print('star vs dc max rel', float(np.max(np.abs(r["area"] - r3["area"]) / r3["area"])))
