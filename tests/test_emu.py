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
