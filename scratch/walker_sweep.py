"""cfg 3 throughput of the walker kernel (device-resident) at several batch sizes; GTA_B200_LPF_SPLIT / _SLOTS / _FSLOTS from the environment."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200.api as api
if os.environ.get('GTA_LIB'): api.lib_path = lambda: os.path.abspath(os.environ['GTA_LIB'])     # A/B against another build
import g_tessla_b200 as g
from g_tessla_b200 import synth
cfg = int(sys.argv[1]); sizes = [int(x) for x in sys.argv[2].split(',')]
base = synth.config(cfg, nframes=2000)
N = base['xyz'].shape[1]
maxp = N + 4 + 4 * int(base['box'][:, :2].max() / 0.8)
want = None
for nf in sizes:
    reps = (nf + 1999) // 2000
    xyz = torch.from_numpy(np.tile(base['xyz'], (reps, 1, 1))[:nf]).cuda(); box = torch.from_numpy(np.tile(base['box'], (reps, 1))[:nf]).cuda()
    T = g.Tessellator(nf, N, maxp, 0.8, base['flags'])
    best = 1e9
    for it in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); T.run(xyz, box, 3); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    a = T.area.cpu().numpy()
    stage = ''
    if hasattr(g.lib(), 'gta_b200_stage_ms'):
        import ctypes as C
        L = g.lib(); L.gta_b200_stage_reset(-1); T.run(xyz, box, 3); torch.cuda.synchronize()
        v = [C.c_double() for _ in range(3)]; nl = C.c_int()
        L.gta_b200_stage_ms(-1, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(nl))
        stage = ' stages(prepass/walk/area ms)=%.1f/%.1f/%.1f x%d' % (v[0].value, v[1].value, v[2].value, nl.value)
    if want is None:
        os.environ.setdefault('X', '1'); L = g.lib(); old = L.gta_b200_set_batch_threshold(0)
        want = g.tessellate(base['xyz'][:300], base['box'][:300], 0.8, base['flags'])['area']; L.gta_b200_set_batch_threshold(old)
    m = min(300, nf)
    print('cfg%d split=%s slots=%s frames %d: %.2f ms -> %.0f frames/s' % (cfg, os.environ.get('GTA_B200_LPF_SPLIT'), os.environ.get('GTA_B200_LPF_SLOTS'), nf, best * 1e3, nf / best),
          'bad', int((T.status != 0).sum()), 'maxrel vs fused', float(np.max(np.abs(a[:m] - want[:m]) / want[:m])), stage, flush=True)
