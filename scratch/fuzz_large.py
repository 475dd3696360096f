"""Random large frames through the star engine's tiles (gta_b200_triangulate_large under PATH_STAR) against the reference:
sizes, aspect ratios, clusters, voids, points along the bounding box.  python scratch/fuzz_large.py [cases] [seed]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
from oracle import RefOracle, PortOracle, ref_available
chk = RefOracle() if ref_available() else PortOracle()
g.set_path(g.PATH_STAR)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = star = 0
t0 = time.time()
for k in range(cases):
    n = int(rng.integers(3600, 150000))
    w, h = 10.0 ** rng.uniform(0, 2.5), 10.0 ** rng.uniform(0, 2.5)
    kind = int(rng.integers(0, 5))
    p = rng.random((n, 2)) * np.array([w, h])
    if kind == 1:      # clusters on a sparse background
        c = rng.random((6, 2)) * np.array([w, h])
        p[: n // 2] = c[rng.integers(0, 6, n // 2)] + rng.normal(0, 0.01 * min(w, h), (n // 2, 2))
    elif kind == 2:    # a void in the middle
        keep = np.hypot(p[:, 0] / w - 0.5, p[:, 1] / h - 0.5) > 0.3
        p = p[keep]
    elif kind == 3:    # points along the bounding box (the hull -corr makes)
        m = int(np.sqrt(len(p)))
        tx, ty = np.linspace(0, w, m), np.linspace(0, h, m)[1:-1]
        p = np.concatenate([p[(p > 0).all(1) & (p[:, 0] < w) & (p[:, 1] < h)], np.stack([tx, 0 * tx], 1), np.stack([tx, 0 * tx + h], 1),
                            np.stack([0 * ty, ty], 1), np.stack([0 * ty + w, ty], 1)])
    elif kind == 4:    # a disc: a long curved hull
        r, a = np.sqrt(rng.random(n)), rng.random(n) * 2 * np.pi
        p = np.stack([w * r * np.cos(a), w * r * np.sin(a)], 1)
    p = np.unique(p, axis=0)
    tri, st, ms = g.triangulate_large(p)
    eng = g.lib().gta_b200_large_last_engine()
    want, _ = chk.triangulate(p)
    ok = st == 0 and np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want))
    star += eng == 1
    bad += not ok
    print('case %d kind %d n %d %gx%g: status %d engine %d %.1f ms %s' % (k, kind, len(p), w, h, st, eng, ms, 'ok' if ok else 'MISMATCH'), flush=True)
print('%d cases, %d through the stars, %d bad, %.0f s' % (cases, star, bad, time.time() - t0))
sys.exit(1 if bad else 0)
