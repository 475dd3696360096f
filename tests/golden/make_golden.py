"""Regenerate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these fixtures -- outputs of
the reference's own dtriangulate / delaunay_tessellate on seeded inputs -- are what pins the
oracle port, the host emulation and the CUDA path on machines without /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import RefOracle, build_oracles  # noqa: E402
from g_tessla_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def lattice(nx, ny, h=0.5):
    gx, gy = np.meshgrid(np.arange(nx) * h, np.arange(ny) * h)
    return np.stack([gx.ravel(), gy.ravel()], 1)   # row-major, SURVEY.md §8a note


def main():
    build_oracles(ref=True)
    ref = RefOracle()
    wrap = RefOracle(wrap=True)
    out = {}
    # 1. raw 2-D point sets through dtriangulate (reference src/delaunay_tri.c:413)
    rng = np.random.Generator(np.random.Philox(key=77))
    sets = {}
    for n in (2, 3, 4, 5, 7, 8, 13, 33, 100, 257):
        sets["rand%d" % n] = rng.random((n, 2))
    sets["rand_f32_300"] = rng.random((300, 2)).astype(np.float32).astype(np.float64) * 25.0
    for nx, ny in ((9, 7), (12, 5), (4, 4), (3, 17), (16, 16)):
        sets["lattice_%dx%d" % (nx, ny)] = lattice(nx, ny)
    sets["collinear_x"] = np.stack([np.arange(9) * 0.25, np.zeros(9)], 1)
    sets["collinear_diag"] = np.stack([np.arange(7) * 0.5, np.arange(7) * 0.5], 1)
    sets["equal_x_cols"] = np.stack([np.repeat(np.arange(4) * 1.0, 5), np.tile(np.arange(5) * 0.375, 4)], 1)
    th = np.arange(24) * (2 * np.pi / 24)
    sets["circle24_center"] = np.concatenate([np.stack([np.cos(th), np.sin(th)], 1), [[0.0, 0.0]]])
    # near-ties in x (0 < |dx| < 1e-12): the reference comparator orders such runs by y
    # (delaunay_tri.c:115-123); happens with float32 atoms at x = 2.0 next to -corr points at 2*0.8+0.4
    nt = rng.random((60, 2)) * 6.4
    nt[5] = (2.0, 3.1); nt[6] = (2 * 0.8 + 0.4, 0.0); nt[7] = (2 * 0.8 + 0.4, 6.4)
    nt[8] = (7 * 0.8 + 0.4, 0.0); nt[9] = (6.0, 1.7); nt[10] = (6.0, 5.9); nt[11] = (7 * 0.8 + 0.4, 6.4)
    nt[12] = (4.0, 2.0); nt[13] = (4.0 + 2 ** -50, 1.0); nt[14] = (4.0 - 2 ** -50, 3.0); nt[15] = (4.0 + 2 ** -49, 0.5)
    assert 0 < nt[8, 0] - 6.0 < 1e-12
    sets["near_tie_x"] = nt
    for name, p in sets.items():
        tri, nv = ref.triangulate(p)
        assert nv == len(p)
        out["pts_" + name] = p
        out["tri_" + name] = tri
    np.savez_compressed(os.path.join(HERE, "dtriangulate.npz"), **out)

    # 2. delaunay_tessellate (reference src/gta_tri.c:80) on the first frames of each config
    tess = {}
    for cfg, nfr in ((1, 24), (2, 24), (3, 4), (5, 4)):
        c = synth.config(cfg, nframes=nfr)
        wrap.counters(reset=True)
        r = wrap.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=1, want_corr=True)
        cnt = wrap.counters()
        tess["cfg%d_nframes" % cfg] = np.int64(nfr)
        tess["cfg%d_area" % cfg] = r["area"]
        tess["cfg%d_areabox" % cfg] = r["areabox"]
        if r["area2d"] is not None:
            tess["cfg%d_area2d" % cfg] = r["area2d"]
        tess["cfg%d_ncorr" % cfg] = r["ncorr"]
        tess["cfg%d_corr" % cfg] = r["corr"]
        tess["cfg%d_counts" % cfg] = np.array([cnt["orient2d"], cnt["orient2d_zero"], cnt["incircle"], cnt["incircle_zero"]])
        # triangles of frame 0 with the appended points (emission order)
        nc = int(r["ncorr"][0])
        p = np.concatenate([c["xyz"][0, :, :2], r["corr"][0, :nc, :2]])
        tri, _ = ref.triangulate(p)
        tess["cfg%d_tri0" % cfg] = tri
        # checksum of the generator output so a numpy change that alters the stream is noticed
        tess["cfg%d_xyz_sum" % cfg] = np.array([c["xyz"].sum(), c["box"].sum()])
    np.savez_compressed(os.path.join(HERE, "tessellate.npz"), **tess)

    # 3. predicate sign vectors: near-degenerate inputs (reference predicates.c:1620, :3226)
    rng = np.random.Generator(np.random.Philox(key=99))
    o = []
    for _ in range(4000):
        a, b = rng.random(2), rng.random(2)
        t = rng.random()
        c = a + t * (b - a)                         # nearly collinear after rounding
        if rng.random() < 0.3:
            c = np.float32(c).astype(np.float64); a = np.float32(a).astype(np.float64); b = np.float32(b).astype(np.float64)
        scale = 2.0 ** rng.integers(-30, 30)
        o.append(np.concatenate([a, b, c]) * scale)
    q = 2.0 ** -10
    for _ in range(1000):                           # exactly collinear dyadic
        a = rng.integers(0, 4096, 2) * q; d = rng.integers(-50, 50, 2) * q
        o.append(np.concatenate([a, a + d * rng.integers(1, 9), a + d * rng.integers(-9, 0)]))
    o = np.array(o)
    ic = []
    for _ in range(3000):
        ctr = rng.random(2) * 10; rad = rng.random() * 3 + 0.1
        th = np.sort(rng.random(3)) * 2 * np.pi
        abc = ctr[None] + rad * np.stack([np.cos(th), np.sin(th)], 1)
        thd = rng.random() * 2 * np.pi
        d = ctr + rad * np.array([np.cos(thd), np.sin(thd)])   # nearly cocircular
        if rng.random() < 0.3:
            abc = np.float32(abc).astype(np.float64); d = np.float32(d).astype(np.float64)
        ic.append(np.concatenate([abc.ravel(), d]))
    for _ in range(1500):                           # exactly cocircular: rectangle corners on a dyadic grid
        x0, y0 = rng.integers(0, 2048, 2) * q; w, h = rng.integers(1, 64, 2) * q
        corners = np.array([[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]])
        k = rng.integers(0, 4)
        corners = np.roll(corners, k, axis=0)
        if rng.random() < 0.5:
            corners = corners[[0, 2, 1, 3]]
        ic.append(corners.ravel())
    ic = np.array(ic)
    np.savez_compressed(os.path.join(HERE, "predicates.npz"), o2d=o, o2d_sign=np.sign(ref.orient2d(o)).astype(np.int32),
                        icc=ic, icc_sign=np.sign(ref.incircle(ic)).astype(np.int32))
    for f in ("dtriangulate.npz", "tessellate.npz", "predicates.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
