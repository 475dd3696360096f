"""A/B: the product library against scratch/libdbg.so (an experiment build) on the same batches."""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import numpy as np, torch
    sys.path.insert(0, '.')
    import g_tessla_b200.api as api
    if sys.argv[2] == 'dbg': api.lib_path = lambda: os.path.abspath('scratch/libdbg.so')
    import g_tessla_b200 as g
    from g_tessla_b200 import synth
    base = synth.config(3, nframes=2000); N = base['xyz'].shape[1]; maxp = N + 4 + 4 * int(base['box'][:, :2].max() / 0.8)
    for nf in [int(x) for x in sys.argv[3].split(',')]:
        reps = (nf + 1999) // 2000
        xyz = torch.from_numpy(np.tile(base['xyz'], (reps, 1, 1))[:nf]).cuda(); box = torch.from_numpy(np.tile(base['box'], (reps, 1))[:nf]).cuda()
        T = g.Tessellator(nf, N, maxp, 0.8, base['flags']); best = 1e9
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); T.run(xyz, box, 3); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        print('%s frames %d: %.1f ms -> %.0f frames/s bad %d' % (sys.argv[2], nf, best * 1e3, nf / best, int((T.status != 0).sum())))
else:
    for which in ('prod', 'dbg'):
        subprocess.run([sys.executable, __file__, 'child', which, sys.argv[1] if len(sys.argv) > 1 else '66000,100000'])
