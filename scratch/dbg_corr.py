import sys, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
from oracle import RefOracle
ref = RefOracle()
c = synth.config(1, nframes=1000)
w = ref.tessellate(c['xyz'], c['box'], c['espace'], c['flags'], want_corr=True)
r = g.tessellate(c['xyz'], c['box'], c['espace'], c['flags'], want_corr=True)
nc = 36
d = np.argwhere(r['corr'][:, :nc] != w['corr'][:, :nc])
print(len(d), 'differences')
for f, s, k in d[:20]:
    print(f, s, k, repr(r['corr'][f, s, k]), repr(w['corr'][f, s, k]), 'box', c['box'][f])
f = d[0][0]
np.save('gpurun_out/dbg_frame.npy', c['xyz'][f]); np.save('gpurun_out/dbg_box.npy', c['box'][f])
np.save('gpurun_out/dbg_gpu_corr.npy', r['corr'][f]); np.save('gpurun_out/dbg_ref_corr.npy', w['corr'][f])
