"""Throughput of the non-headline BASELINE configs (reported in DESIGN.md; bench.py measures cfg 3)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
from oracle import RefOracle
ref = RefOracle()
out = {}
for cfg, nfr in ((1, 1000), (2, 1000), (1, 20000), (2, 20000), (5, 10000)):
    base = synth.config(cfg, nframes=2000)
    reps = (nfr + 1999) // 2000
    xyz = np.tile(base['xyz'], (reps, 1, 1))[:nfr]; box = np.tile(base['box'], (reps, 1))[:nfr]
    for _ in range(2): g.tessellate(xyz, box, 0.8, base['flags'])   # (sizes the pinned staging buffers and the workspaces)
    t0 = time.perf_counter(); r = g.tessellate(xyz, box, 0.8, base['flags']); t1 = time.perf_counter()
    assert (r['status'] == 0).all()
    w = ref.tessellate(base['xyz'][:1000], base['box'][:1000], 0.8, base['flags'], nthreads=os.cpu_count())
    w1 = ref.tessellate(base['xyz'][:200], base['box'][:200], 0.8, base['flags'], nthreads=1)
    out['cfg%d_host_%d' % (cfg, nfr)] = dict(frames=nfr, gpu_kernel_fps=nfr / (r['kernel_ms'] / 1e3), gpu_e2e_fps=nfr / (t1 - t0),
                              cpu_all_threads_fps=1000 / w['seconds'], cpu_1thread_fps=200 / w1['seconds'], cpu_threads=os.cpu_count(),
                              max_rel=float(np.max(np.abs(r['area'][:1000] - w['area']) / w['area'])))
# device-resident batches (inputs in HBM): the sizes BASELINE.json names and large batches
for cfg, nfr in ((1, 1000), (2, 1000), (1, 100000), (2, 100000), (3, 3000), (3, 12500), (5, 10000), (5, 50000)):
    base = synth.config(cfg, nframes=min(nfr, 2000)); reps = max(1, (nfr + 1999) // 2000)
    N = base['xyz'].shape[1]; maxp = N + 4 + 4 * int(base['box'][:, :2].max() / 0.8 + 1)
    xyz = torch.from_numpy(np.tile(base['xyz'], (reps, 1, 1))[:nfr]).cuda(); box = torch.from_numpy(np.tile(base['box'], (reps, 1))[:nfr]).cuda()
    T = g.Tessellator(len(xyz), N, maxp, 0.8, base['flags'] | (2 if cfg == 2 else 0)); best = 1e9
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); T.run(xyz, box, 3); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    assert int((T.status != 0).sum()) == 0
    cert, fb = g.star_counters(reset=True)
    out['cfg%d_resident_%d' % (cfg, len(xyz))] = dict(frames=len(xyz), ms=best * 1e3, fps=len(xyz) / best,
                                                      engine='star engine' if fb == 0 and cert > 0 else ('divide and conquer (tied frames handed over / stars skipped)'))
p = synth.uniform_patch(1_000_000, seed=4)
g.triangulate_large(p[:100000])
tri, st, ms = g.triangulate_large(p)
t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
t2 = time.perf_counter(); want, _ = ref.triangulate(p); t3 = time.perf_counter()
out['cfg4_divide_and_conquer'] = dict(points=len(p), triangles=len(tri), status=st, gpu_device_ms=ms, gpu_e2e_ms=(t1 - t0) * 1e3, cpu_1thread_ms=(t3 - t2) * 1e3,
                                      bit_exact_in_emission_order=bool(np.array_equal(tri, want)))
g.set_path(g.PATH_STAR)
tri, st, ms = g.triangulate_large(p)
t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
out['cfg4_star_engine'] = dict(points=len(p), triangles=len(tri), status=st, engine=int(g.lib().gta_b200_large_last_engine()), gpu_device_ms=ms, gpu_e2e_ms=(t1 - t0) * 1e3,
                               set_equal_after_canonical_sorting=bool(np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want))))
g.set_path(g.PATH_AUTO)
print(json.dumps(out, indent=1))
