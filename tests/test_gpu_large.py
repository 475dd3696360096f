"""GPU: the single-large-frame path (global memory, 32-bit indices) against the oracle, and
size-independent properties at BASELINE cfg 4's full size (10^6 points)."""
import numpy as np
import pytest

from g_tessla_b200 import synth

pytestmark = pytest.mark.gpu


def test_large_path_matches_oracle_small_and_medium(gta, checker):
    rng = np.random.default_rng(4)
    for n in (2, 3, 7, 100, 1156, 5000, 40000):
        p = rng.random((n, 2)) * 50
        tri, st, _ = gta.triangulate_large(p)
        assert st == 0, (n, st)
        want, _ = checker.triangulate(p)
        assert np.array_equal(tri, want), n                  # same triangles in the same emission order


def test_large_path_degenerate_inputs(gta, checker, golden):
    g = golden["dtriangulate"]
    for k in g.files:
        if k.startswith("pts_") and len(g[k]) >= 2:
            tri, st, _ = gta.triangulate_large(g[k])
            assert st == 0, k
            assert np.array_equal(tri, g["tri_" + k[4:]]), k
    # a 200 x 150 exactly cocircular lattice: the diagonal pattern follows the recursion tree
    gx, gy = np.meshgrid(np.arange(200) * 0.25, np.arange(150) * 0.25)
    p = np.stack([gx.ravel(), gy.ravel()], 1)
    tri, st, _ = gta.triangulate_large(p)
    want, _ = checker.triangulate(p)
    assert st == 0 and np.array_equal(tri, want)


def test_large_path_status(gta):
    p = np.random.default_rng(1).random((5000, 2)); p[77] = p[4001]
    _, st, _ = gta.triangulate_large(p)
    assert st & 2
    p[77, 0] = np.inf
    _, st, _ = gta.triangulate_large(p)
    assert st & 64


def test_cfg4_million_points_properties(gta):
    """cfg 4: 10^6 uniform points (float-rounded) in an 800 x 800 nm patch. The oracle takes ~6 s per
    frame on one core, so this test checks properties; test_cfg4_vs_oracle compares a full frame."""
    from scipy.spatial import ConvexHull
    p = synth.uniform_patch(1_000_000, seed=4)
    tri, st, ms = gta.triangulate_large(p)
    assert st == 0
    hull = len(ConvexHull(p).vertices)
    # ntri = 2(n-1) - k with k hull vertices incl. collinear ones (reference delaunay_tri.c:608);
    # qhull drops collinear hull points, so k >= its count
    assert len(tri) <= 2 * (len(p) - 1) - hull and len(tri) >= 2 * (len(p) - 1) - hull - 64
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    cross = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    assert (cross > 0).all()                                   # every triangle counter-clockwise
    assert abs(0.5 * cross.sum() - ConvexHull(p).volume) <= 1e-9 * ConvexHull(p).volume   # tiles the hull exactly once
    assert (tri[:, 0] != tri[:, 1]).all() and len(np.unique(np.sort(tri, 1), axis=0)) == len(tri)


def test_cfg4_vs_oracle(gta, checker):
    p = synth.uniform_patch(1_000_000, seed=4)
    tri, st, ms = gta.triangulate_large(p)
    want, _ = checker.triangulate(p)
    assert st == 0
    assert np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want))
    assert np.array_equal(tri, want)
