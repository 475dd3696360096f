"""CPU: the oracle port against the reference's outputs (golden fixtures + live oracle/_ref)."""
import numpy as np
import pytest

from g_tessla_b200 import synth


def _names(g):
    return sorted(k[4:] for k in g.files if k.startswith("pts_"))


def test_port_matches_golden_triangulations(port, golden):
    g = golden["dtriangulate"]
    for name in _names(g):
        tri, nv = port.triangulate(g["pts_" + name])
        assert nv == len(g["pts_" + name])
        assert np.array_equal(tri, g["tri_" + name]), name      # same triangles, same emission order


def test_lattice_tiebreak_pattern(port, golden):
    """Known answer observed from the reference (SURVEY.md §8a note): on the exactly cocircular
    9x7 lattice the diagonal choice follows the D&C recursion tree."""
    g = golden["dtriangulate"]
    nx, ny = 9, 7
    tri, _ = port.triangulate(g["pts_lattice_9x7"])
    rows = []
    for j in range(ny - 2, -1, -1):          # top row first
        s = ""
        for i in range(nx - 1):
            a, b, c, d = j * nx + i, j * nx + i + 1, (j + 1) * nx + i + 1, (j + 1) * nx + i
            has_ac = any({a, c} <= set(t) for t in tri.tolist())
            has_bd = any({b, d} <= set(t) for t in tri.tolist())
            assert has_ac != has_bd
            s += "/" if has_ac else "\\"
        rows.append(s)
    assert rows == ["/////\\\\/", "////\\\\//", "///\\\\///", "//\\\\////", "/\\\\/////", "\\\\//////"]


@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_port_matches_golden_tessellate(port, golden, cfg):
    t = golden["tessellate"]
    nfr = int(t["cfg%d_nframes" % cfg])
    c = synth.config(cfg, nframes=nfr)
    assert np.array_equal(np.array([c["xyz"].sum(), c["box"].sum()]), t["cfg%d_xyz_sum" % cfg]), "generator stream changed"
    r = port.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
    assert np.array_equal(r["ncorr"], t["cfg%d_ncorr" % cfg])
    assert np.array_equal(r["corr"], t["cfg%d_corr" % cfg])               # appended points bit-exact
    assert np.array_equal(r["areabox"], t["cfg%d_areabox" % cfg])
    assert np.array_equal(r["area"], t["cfg%d_area" % cfg])               # same summation order => bit-exact
    if cfg == 2:
        assert np.array_equal(r["area2d"], t["cfg2_area2d"])
        # README.md:16 of the reference: corrected 2-D area == box area
        assert np.allclose(r["area2d"], r["areabox"], rtol=1e-12, atol=0)
    nc = int(r["ncorr"][0])
    p = np.concatenate([c["xyz"][0, :, :2], r["corr"][0, :nc, :2]])
    tri, _ = port.triangulate(p)
    assert np.array_equal(tri, t["cfg%d_tri0" % cfg])


def test_port_predicate_signs_golden(port, golden):
    g = golden["predicates"]
    assert np.array_equal(np.sign(port.orient2d(g["o2d"])).astype(np.int32), g["o2d_sign"])
    assert np.array_equal(np.sign(port.incircle(g["icc"])).astype(np.int32), g["icc_sign"])


def test_predicate_signs_exact_rational(port):
    """Independent pin: exact rational arithmetic on near-degenerate inputs."""
    from fractions import Fraction as Fr
    rng = np.random.default_rng(5)
    o, ic = [], []
    for _ in range(300):
        a, b = rng.random(2), rng.random(2)
        o.append(np.concatenate([a, b, a + rng.random() * (b - a)]))
        ctr = rng.random(2); th = rng.random(4) * 6.283
        ic.append((ctr[None] + np.stack([np.cos(th), np.sin(th)], 1)).ravel())
    o, ic = np.array(o), np.array(ic)
    so, si = np.sign(port.orient2d(o)), np.sign(port.incircle(ic))
    for k in range(len(o)):
        ax, ay, bx, by, cx, cy = map(Fr, o[k].tolist())
        d = (ax - cx) * (by - cy) - (ay - cy) * (bx - cx)
        assert so[k] == (d > 0) - (d < 0)
        ax, ay, bx, by, cx, cy, dx, dy = map(Fr, ic[k].tolist())
        adx, ady, bdx, bdy, cdx, cdy = ax - dx, ay - dy, bx - dx, by - dy, cx - dx, cy - dy
        d = ((adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) + (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy)
             + (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady))
        assert si[k] == (d > 0) - (d < 0)


def test_port_vs_live_reference(port, ref):
    """Where oracle/_ref exists: port == unmodified reference on fresh seeded inputs."""
    rng = np.random.default_rng(11)
    for n in (2, 3, 6, 50, 400, 1156):
        p = rng.random((n, 2)) * 20
        a, _ = ref.triangulate(p)
        b, _ = port.triangulate(p)
        assert np.array_equal(a, b)
    for cfg in (1, 2, 5):
        c = synth.config(cfg, nframes=40)
        ra = ref.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=2, want_corr=True)
        rb = port.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=2, want_corr=True)
        for k in ("area", "areabox", "corr", "ncorr"):
            assert np.array_equal(ra[k], rb[k]), (cfg, k)
        if cfg == 2:
            assert np.array_equal(ra["area2d"], rb["area2d"])


def test_triangle_count_formula(port):
    """ntri = 2(n-1) - k, k = hull vertices (reference delaunay_tri.c:608)."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(2)
    for n in (10, 100, 1000):
        p = rng.random((n, 2))
        tri, _ = port.triangulate(p)
        assert len(tri) == 2 * (n - 1) - len(ConvexHull(p).vertices)


def test_matches_qhull_on_general_position(port):
    from scipy.spatial import Delaunay
    p = np.random.default_rng(3).random((1024, 2))
    tri, _ = port.triangulate(p)
    assert np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(Delaunay(p).simplices))
