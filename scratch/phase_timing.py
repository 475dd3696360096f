"""Per-phase cycle breakdown of the batched kernel (needs a -DGT_PHASE_TIMING build: scratch/libdbg.so)."""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200.api as api
api.lib_path = lambda: os.path.abspath('scratch/libdbg.so')
from g_tessla_b200 import synth
L = api.lib()
out = (C.c_ulonglong * 64)()
L.gta_b200_dbg_timers(None, 1)
c = synth.config(3, nframes=3000)
api.tessellate(c['xyz'], c['box'], c['espace'], c['flags'])
L.gta_b200_dbg_timers(None, 1)
r = api.tessellate(c['xyz'], c['box'], c['espace'], c['flags'])
L.gta_b200_dbg_timers(out, 0)
v = np.array(list(out), dtype=np.float64); nf = v[21]
names = ['load', 'corr', 'sort', 'postsort'] + ['level%d' % d for d in range(16)] + ['enum+area']
tot = v[:21].sum()
print('coop_nodes', os.environ.get('GTA_B200_COOP_NODES'), 'frames', nf, 'kernel_ms', r['kernel_ms'], 'cycles/frame %.0f' % (tot / nf))
for i, nm in enumerate(names):
    if v[i]: print('%-10s %9.0f cyc/frame %5.1f%%' % (nm, v[i] / nf, 100 * v[i] / tot))
if v[26]: print('coop merges %d: tangent+first %.0f cyc/merge, loop %.0f cyc/step, rounds/step %.2f, steps/merge %.1f' % (v[26], v[22]/v[26], v[23]/v[25], v[24]/v[25], v[25]/v[26]))
