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
