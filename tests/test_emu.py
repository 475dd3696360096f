"""CPU: the device headers (dt_core.cuh, gt_predicates.cuh) compiled for the host and driven in the
kernel's order, against the golden fixtures and the oracle. Test infrastructure only."""
import numpy as np

from g_tessla_b200 import synth


def test_emu_matches_golden_triangulations(emu, golden):
    g = golden["dtriangulate"]
    for k in g.files:
        if k.startswith("pts_"):
            tri = emu.triangulate(g[k])
            assert np.array_equal(tri, g["tri_" + k[4:]]), k


def test_emu_matches_oracle_on_configs(emu, checker):
    for cfg, nfr in ((1, 30), (3, 6), (5, 6)):
        c = synth.config(cfg, nframes=nfr)
        r = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
        for f in range(nfr):
            nc = int(r["ncorr"][f])
            p = np.concatenate([c["xyz"][f, :, :2], r["corr"][f, :nc, :2]])
            a, _ = checker.triangulate(p)
            assert np.array_equal(emu.triangulate(p), a), (cfg, f)


def test_emu_random_sizes(emu, checker):
    rng = np.random.default_rng(0)
    for n in list(range(2, 40)) + [63, 64, 65, 127, 128, 129, 500, 1155, 1157, 2048, 3000]:
        p = rng.random((n, 2))
        a, _ = checker.triangulate(p)
        assert np.array_equal(emu.triangulate(p), a), n


def test_device_predicates_signs(emu, golden):
    g = golden["predicates"]
    assert np.array_equal(emu.signs("emu_orient2d_signs", g["o2d"], 6), g["o2d_sign"])
    assert np.array_equal(emu.signs("emu_incircle_signs", g["icc"], 8), g["icc_sign"])
    # the exact stage alone must agree everywhere, not only where the filter is inconclusive
    assert np.array_equal(emu.signs("emu_orient2d_exact", g["o2d"], 6), g["o2d_sign"])
    assert np.array_equal(emu.signs("emu_incircle_exact", g["icc"], 8), g["icc_sign"])


def test_exact_stage_wide_exponent_range(emu, port):
    """Operands spanning many binades select the 2- and 4-word integer paths."""
    rng = np.random.default_rng(8)
    o = rng.random((2000, 6)) * (2.0 ** rng.integers(-20, 20, size=(2000, 6)))
    ic = rng.random((2000, 8)) * (2.0 ** rng.integers(-12, 12, size=(2000, 8)))
    so = emu.signs("emu_orient2d_exact", o, 6)
    si = emu.signs("emu_incircle_exact", ic, 8)
    ok_o, ok_i = so != 2, si != 2
    assert ok_o.mean() > 0.9 and ok_i.mean() > 0.9
    assert np.array_equal(so[ok_o], np.sign(port.orient2d(o))[ok_o].astype(np.int32))
    assert np.array_equal(si[ok_i], np.sign(port.incircle(ic))[ok_i].astype(np.int32))


def test_walkers_any_split_any_interleaving(emu, checker, golden):
    """The throughput kernel cuts a frame's recursion tree into 2^level subtrees whose walkers run concurrently, share the
    frame's edge pool and hand over to their parents through the "arrived" bits (lpf_core.cuh::lp_finish_node). Whatever
    the cut and the interleaving, the triangles must be the reference's, in its emission order, and the pool must not
    grow beyond the final edge count 3n - 3 - hull (no edge is lost between walkers)."""
    g = golden["dtriangulate"]
    for k in g.files:
        if k.startswith("pts_") and len(g[k]) >= 2:
            for level, seed in ((0, 0), (2, 5), (5, 9)):
                tri, _ = emu.triangulate_walkers(g[k], level, seed)
                assert np.array_equal(tri, g["tri_" + k[4:]]), (k, level, seed)
    for cfg, nfr in ((3, 3), (5, 3), (1, 6)):
        c = synth.config(cfg, nframes=nfr)
        r = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
        for f in range(nfr):
            nc = int(r["ncorr"][f])
            p = np.concatenate([c["xyz"][f, :, :2], r["corr"][f, :nc, :2]])
            want, _ = checker.triangulate(p)
            edges = (3 * len(want) + nc) // 2 if cfg != 5 else None      # every -corr point is on the hull: E = (3T + h) / 2
            for level in range(6):
                for seed in (0, 3, 11):
                    tri, st = emu.triangulate_walkers(p, level, seed)
                    assert np.array_equal(tri, want), (cfg, f, level, seed)
                    if edges is not None:
                        assert st[1] == edges, (cfg, f, level, st)
            # the root walker's iteration count is the frame's latency in scheduler rounds: it must shrink with the cut
            it0 = emu.triangulate_walkers(p, 0, 1)[1][3]; it4 = emu.triangulate_walkers(p, 4, 1)[1][3]
            assert it4 * 3 < it0
    rng = np.random.default_rng(4)
    for n in (2, 3, 4, 5, 7, 8, 9, 16, 33, 64, 65, 200):
        p = rng.random((n, 2))
        want, _ = checker.triangulate(p)
        for level in (0, 1, 3, 5):
            assert np.array_equal(emu.triangulate_walkers(p, level, 7)[0], want), (n, level)


def test_cfg5_degeneracy_census(golden):
    """the cfg 5 generator must be a real tie-break stress -- the reference run that produced the golden
    fixture evaluated thousands of EXACTLY ZERO predicates (counted by the --wrap build, tests/golden/make_golden.py),
    several times what the generic cfg 3 frames give (whose zeros are the collinear -corr boundary and the reference's
    incircle calls with a repeated point)."""
    o2d, o2d_zero, icc, icc_zero = (int(v) for v in golden["tessellate"]["cfg5_counts"])
    nfr = int(golden["tessellate"]["cfg5_nframes"])
    assert o2d_zero >= 2000 * nfr and icc_zero >= 300 * nfr, (o2d_zero, icc_zero, nfr)
    g3 = golden["tessellate"]["cfg3_counts"]; n3 = int(golden["tessellate"]["cfg3_nframes"])
    assert o2d_zero / nfr > 5 * int(g3[1]) / n3


def test_cfg5_census_live(ref, tmp_path):
    """The same census against the live reference where oracle/_ref exists (a generator change cannot hide behind the
    fixture), and the part of it that is the point of cfg 5: exactly cocircular quadruples of four DISTINCT points."""
    import os
    import pytest
    from oracle import RefOracle
    from oracle.loader import REF_WRAP_SO
    if not os.path.exists(REF_WRAP_SO):
        pytest.skip("counter build of the reference not available")
    w = RefOracle(wrap=True)
    out = {}
    for cfg in (5, 3):
        c = synth.config(cfg, nframes=4)
        dump = str(tmp_path / ("calls%d.bin" % cfg))
        w.counters(reset=True); w.set_dump(dump)
        w.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=1)
        w.set_dump(None)
        cnt = w.counters()
        rec = np.fromfile(dump, dtype=np.float64).reshape(-1, 10)
        ic = rec[rec[:, 0] == 3.0]
        pts = ic[:, 1:9].reshape(-1, 4, 2)
        distinct = np.ones(len(ic), bool)
        for i in range(4):
            for j in range(i + 1, 4):
                distinct &= ~np.all(pts[:, i] == pts[:, j], axis=1)
        out[cfg] = (cnt, int(np.sum((ic[:, 9] == 0.0) & distinct)))
    cnt5, cocirc5 = out[5]; cnt3, cocirc3 = out[3]
    assert cnt5["orient2d_zero"] >= 8000 and cnt5["incircle_zero"] >= 1200, cnt5
    assert cocirc5 >= 200, cocirc5            # >= 50 exactly cocircular decisions per frame: the tie-break rule decides them
    assert cocirc3 == 0, cocirc3              # generic positions: none


def _canon(t):
    return synth.canonical_triangles(t)


def test_star_path_matches_oracle_on_configs(emu, checker):
    """star_core.cuh on the host: for frames whose Delaunay triangulation is unique the stars deliver the reference's
    triangle set (canonical order); the lattice of cfg 5 (cocircular ties everywhere) must NOT be certified."""
    for cfg, nfr in ((1, 30), (2, 10), (3, 6)):
        c = synth.config(cfg, nframes=nfr)
        r = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
        for f in range(nfr):
            p = np.concatenate([c["xyz"][f, :, :2], r["corr"][f, :int(r["ncorr"][f]), :2]])
            want, _ = checker.triangulate(p)
            for full, ccap in ((True, 64), (False, 64), (False, 12)):
                tri, st, _ = emu.star(p, full=full, ccap=ccap)
                assert st == 0 and np.array_equal(_canon(tri), _canon(want)), (cfg, f, st, full, ccap)
    c = synth.config(5, nframes=3)
    r = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
    for f in range(3):
        p = np.concatenate([c["xyz"][f, :, :2], r["corr"][f, :int(r["ncorr"][f]), :2]])
        tri, st, _ = emu.star(p)
        assert tri is None and st & 1


def test_star_path_random_sizes_and_windows(emu, checker):
    """No -corr points: the hull is irregular, hull vertices scan the whole grid. Any first window and any cell density must
    give the same triangles (the window only has to be grown more or less often)."""
    rng = np.random.default_rng(5)
    for n in list(range(2, 30)) + [64, 65, 129, 500, 1157, 3000]:
        p = rng.random((n, 2)) * np.array([3.0, 7.0])
        want, _ = checker.triangulate(p)
        for w0, dens, full, ccap in ((2, 1.0, True, 64), (0, 1.0, True, 12), (1, 4.0, False, 64), (3, 0.5, False, 12), (2, 1.0, False, 12)):
            tri, st, _ = emu.star(p, w0, dens, full, ccap)
            assert st == 0 and np.array_equal(_canon(tri), _canon(want)), (n, w0, dens, full, ccap, st)
    # clustered points: windows must grow a lot, the result stays exact
    p = np.concatenate([rng.normal(0, 0.01, (300, 2)), rng.normal(5, 2.0, (300, 2)), rng.random((50, 2)) * 40])
    want, _ = checker.triangulate(p)
    for full in (True, False):
        tri, st, cnt = emu.star(p, full=full)
        assert st == 0 and cnt[4] > 0 and np.array_equal(_canon(tri), _canon(want))


def test_star_path_refuses_what_it_cannot_certify(emu, golden):
    """Cocircular lattices, collinear sets and duplicates are left to the divide-and-conquer path (status bits, no triangles);
    the golden near-degenerate sets it does certify must equal the reference's triangles."""
    gx, gy = np.meshgrid(np.arange(9.0), np.arange(7.0))
    tri, st, _ = emu.star(np.stack([gx.ravel(), gy.ravel()], 1))
    assert tri is None and st & 1                                        # squares: four cocircular points on empty circles
    tri, st, _ = emu.star(np.stack([np.arange(10.0), 2 * np.arange(10.0)], 1))
    assert tri is None                                                   # all collinear: the triangle count rule fails
    p = np.random.default_rng(2).random((40, 2)); p[17] = p[3]
    tri, st, _ = emu.star(p)
    assert tri is None and st & 2
    g = golden["dtriangulate"]
    seen = 0
    for k in g.files:
        if k.startswith("pts_") and len(g[k]) >= 3:
            tri, st, _ = emu.star(g[k])
            if tri is not None:
                seen += 1
                assert np.array_equal(_canon(tri), _canon(g["tri_" + k[4:]])), k
    assert seen >= 3


def test_star_path_scales_offsets_and_near_ties(emu, checker):
    """The float screening inside the star path must never change a result: point sets at small and huge scales, far from
    the origin (where the float copies of the coordinates cannot tell neighbours apart and the screen has to stand back),
    and lattices jittered by less than float resolution (every square nearly cocircular: double / exact stages decide).
    A certified result equals the reference's triangles; exact lattices are still refused."""
    rng = np.random.default_rng(11)
    for n in (40, 300):
        base = rng.random((n, 2)) * np.array([5.0, 3.0])
        for scale in (1e-8, 1e-3, 1.0, 1e6, 1e20):       # (below ~1e-10 the reference calls neighbours duplicates: 1e-12 absolute)
            for off in (0.0, 1e3, 3e7):
                p = (base + off) * scale
                want, _ = checker.triangulate(p)
                for full in (True, False):
                    tri, st, cnt = emu.star(p, full=full)
                    assert st == 0 and np.array_equal(_canon(tri), _canon(want)), (n, scale, off, full, st)
    gx, gy = np.meshgrid(np.arange(12.0), np.arange(10.0))
    lat = np.stack([gx.ravel(), gy.ravel()], 1)
    for jit in (1e-3, 1e-7, 1e-9, 1e-13):
        for off in (0.0, 1000.0):
            p = lat + off + rng.uniform(-jit, jit, lat.shape)
            want, _ = checker.triangulate(p)
            tri, st, _ = emu.star(p)
            if tri is None:
                assert st & 1                                            # (a jitter that rounds away leaves exact ties)
            else:
                assert np.array_equal(_canon(tri), _canon(want)), (jit, off)
    tri, st, _ = emu.star(lat * 0.1 + 7.0)
    assert tri is None and st & 1


def test_star_tiles_defer_what_they_cannot_see(emu, checker):
    """Tiles of the grid with a halo (StarGrid GL = 2, the large frame's shared-memory kernel): a tile's own points build their
    stars from the staged cells and DEFER whenever a search would leave them on an open side; tiles + deferred points together
    must give exactly the reference's triangles, for any tile size and halo (small ones defer nearly everything along their
    borders), on uniform, clustered and anisotropic sets."""
    rng = np.random.default_rng(21)
    sets = [rng.random((n, 2)) * np.array([4.0, 3.0]) for n in (50, 700, 3000)]
    sets.append(np.concatenate([rng.normal(0, 0.02, (400, 2)), rng.normal(3, 1.0, (400, 2)), rng.random((60, 2)) * 20]))
    sets.append(rng.random((1500, 2)) * np.array([30.0, 1.0]))
    deferred = tiles = 0
    for p in sets:
        want, _ = checker.triangulate(p)
        for tw, halo, dens in ((32, 6, 1.0), (8, 6, 1.0), (4, 2, 1.0), (5, 3, 0.5), (16, 4, 2.0), (3, 0, 1.0)):
            tri, st, cnt = emu.star_tiles(p, tw, halo, dens)
            assert st == 0 and np.array_equal(_canon(tri), _canon(want)), (len(p), tw, halo, dens, st)
            deferred += cnt[0]; tiles += cnt[1]
    assert deferred > 0 and tiles > len(sets) * 6                        # (both ways were taken)
    # a lattice is refused whichever way its points are done
    gx, gy = np.meshgrid(np.arange(20.0), np.arange(17.0))
    tri, st, _ = emu.star_tiles(np.stack([gx.ravel(), gy.ravel()], 1), 6, 3)
    assert tri is None and st & 1
