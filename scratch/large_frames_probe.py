import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
G, F = 316, 48
xyz, box = synth.jittered_bilayer(F, G, seed=5)
out = []
for rep in range(2):
    for f in range(F):
        b9 = np.zeros((3, 3)); b9[0, 0], b9[1, 1], b9[2, 2] = box[f]
        t0 = time.perf_counter(); r = g.tessellate_large(xyz[f], b9, 0.8, flags=3); t1 = time.perf_counter()
        if rep: out.append((f, r['engine'], r['status'], r['ncorr'], (t1 - t0) * 1e3))
ms = np.array([o[4] for o in out])
print('engines', sorted(set(o[1] for o in out)), 'mean %.2f ms, median %.2f, max %.2f' % (ms.mean(), np.median(ms), ms.max()))
for o in sorted(out, key=lambda o: -o[4])[:6]: print('frame %d engine %d status %d ncorr %d: %.2f ms' % o)
