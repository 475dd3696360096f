"""Per-phase cycles of the lane-per-frame kernel (needs a -DGT_PHASE_TIMING build: scratch/libdbg.so)."""
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
v = np.array(list(out), dtype=np.float64)[32:]
names = ['LEAF', 'TINIT', 'TAN', 'INS', 'SCAN', 'V', 'C', 'fetch']
tot = sum(v[3 * i] for i in range(8))
print('frames', nf, 'warp-cycles in steps %.3g, barrier wait %.3g (%.0f per warp-round), warp-rounds %.3g' % (tot, v[30], v[30] / max(v[31], 1), v[31]))
print('rounds per block %.0f, live walkers per round and block %.1f' % (v[28] / 148, v[27] / max(v[28], 1)))
for i, nm in enumerate(names):
    cyc, ch, ln = v[3 * i], v[3 * i + 1], v[3 * i + 2]
    if ch: print('%-6s %5.1f%% of step cycles  %7.0f cycles/chunk  %5.1f lanes/chunk  %8.0f chunks/frame-equivalent %.0f lane-iters/frame' % (nm, 100 * cyc / tot, cyc / ch, ln / ch, ch * 32 / nf, ln / nf))
