"""Per-section cycles of the pre-pass kernel (needs a -DGT_PHASE_TIMING build: scratch/libdbg.so)."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200.api as api
api.lib_path = lambda: os.path.abspath('scratch/libdbg.so')
import g_tessla_b200 as g
from g_tessla_b200 import synth
L = api.lib()
nf = int(sys.argv[1]); cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out = (C.c_ulonglong * 64)()
L.gta_b200_dbg_timers(None, 1)
c = synth.config(cfg, nframes=2000); reps = (nf + 1999) // 2000
N = c['xyz'].shape[1]; maxp = N + 4 + 4 * int(c['box'][:, :2].max() / 0.8)
xyz = torch.from_numpy(np.tile(c['xyz'], (reps, 1, 1))[:nf]).cuda(); box = torch.from_numpy(np.tile(c['box'], (reps, 1))[:nf]).cuda()
T = g.Tessellator(nf, N, maxp, 0.8, c['flags'])
T.run(xyz, box, 3); torch.cuda.synchronize()
L.gta_b200_dbg_timers(None, 1)
T.run(xyz, box, 3); torch.cuda.synchronize()
L.gta_b200_dbg_timers(out, 0)
v = np.array(list(out), dtype=np.float64)
names = ['tile wait', 'finite check', '-corr', 'minmax+hist', 'scan', 'scatter+bucket order', 'gather', 'near-tie+dups', 'records']
fr = max(v[9], 1)
print('pre-pass: %d frames, %.0f cycles per frame (one CTA)' % (fr, sum(v[:9]) / fr))
for i, nm in enumerate(names):
    print('  %-22s %8.0f cycles  %5.1f%%' % (nm, v[i] / fr, 100 * v[i] / max(sum(v[:9]), 1)))
