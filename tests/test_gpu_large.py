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


def test_large_star_engine_vs_oracle(gta, checker, star_path):
    """The star engine on one large frame (csrc/gta_large_star.cuh; triangle lists only under PATH_STAR: they come in cell
    order): the reference's triangle set after canonical sorting, for point sets with irregular hulls of any size."""
    L = gta.lib()
    rng = np.random.default_rng(4)
    for n in (3, 7, 100, 1156, 5000, 40000, 300000):
        p = rng.random((n, 2)) * np.array([50.0, 20.0])
        tri, st, _ = gta.triangulate_large(p)
        assert st == 0 and L.gta_b200_large_last_engine() == 1, (n, st)
        want, _ = checker.triangulate(p)
        assert np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want)), n
    # a lattice is all ties: not certified, the level-by-level path answers in the reference's emission order
    gx, gy = np.meshgrid(np.arange(60) * 0.25, np.arange(40) * 0.25)
    p = np.stack([gx.ravel(), gy.ravel()], 1)
    tri, st, _ = gta.triangulate_large(p)
    want, _ = checker.triangulate(p)
    assert st == 0 and L.gta_b200_large_last_engine() == 0 and np.array_equal(tri, want)


def test_large_star_tiles_clusters_strips_and_box_hulls(gta, checker, star_path):
    """The tiles of the large frame's grid (ls_tile_kernel) on what they handle least gracefully: clusters that overflow a
    tile's arrays or leave tiles nearly empty (whole tiles deferred), long thin strips (tiles cut by the grid's sides, hull
    everywhere), and a hull that is the bounding box with points along it (what -corr makes of a frame). Tiles and deferred
    points together must give the reference's triangle set."""
    L = gta.lib()
    rng = np.random.default_rng(17)
    sets = [
        np.concatenate([rng.normal(0, 0.01, (8000, 2)), rng.normal(40, 3.0, (30000, 2)), rng.random((2000, 2)) * 100]),
        rng.random((50000, 2)) * np.array([2000.0, 3.0]),
        rng.random((50000, 2)) * np.array([2.0, 900.0]),
    ]
    q = rng.random((90000, 2)) * np.array([30.0, 30.0])
    t = np.linspace(0.0, 30.0, 301)
    sets.append(np.concatenate([q[(q > 0).all(1) & (q < 30).all(1)], np.stack([t, 0 * t], 1), np.stack([t, 0 * t + 30.0], 1),
                                np.stack([0 * t[1:-1], t[1:-1]], 1), np.stack([0 * t[1:-1] + 30.0, t[1:-1]], 1)]))
    # a blob five times as dense as its surroundings: its tiles hold more points than their arrays and are deferred whole
    sets.append(np.concatenate([rng.random((40000, 2)) * 100.0, 50.0 + 8.0 * rng.standard_normal((6000, 2))]))
    for k, p in enumerate(sets):
        tri, st, _ = gta.triangulate_large(p)
        # (set 0 is so clustered -- most points in cells of > 24 -- that the grid is not used at all: level-by-level path)
        assert st == 0 and L.gta_b200_large_last_engine() == (0 if k == 0 else 1), (k, st)
        want, _ = checker.triangulate(p)
        assert np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want)), k


def test_cfg4_star_engine(gta, checker, star_path):
    """cfg 4 through the stars: 10^6 points, the reference's triangle set bit-exact after canonical sorting (north_star)."""
    p = synth.uniform_patch(1_000_000, seed=4)
    tri, st, ms = gta.triangulate_large(p)
    assert st == 0 and gta.lib().gta_b200_large_last_engine() == 1
    want, _ = checker.triangulate(p)
    assert np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want))
    print("cfg 4, star engine: %.1f ms on the device" % ms)


def test_large_surface_area_both_engines(gta):
    """delaunay_surface_area's engine for frames beyond the shared-memory kernels: the star engine by default (areas have no
    order contract), the level-by-level path under PATH_DC; same count, areas equal to the rounding of the sum order."""
    rng = np.random.default_rng(9)
    xyz = np.concatenate([rng.random((200000, 2)) * 300.0, 4.0 + 0.2 * rng.standard_normal((200000, 1))], 1)
    a = gta.surface_area_large(xyz, flags=synth.GTA_2D)
    old = gta.set_path(gta.PATH_DC)
    try:
        b = gta.surface_area_large(xyz, flags=synth.GTA_2D)
    finally:
        gta.set_path(old)
    assert a["status"] == 0 and b["status"] == 0 and a["engine"] == 1 and b["engine"] == 0
    assert a["ntri"] == b["ntri"]
    assert abs(a["area"] - b["area"]) <= 1e-12 * b["area"] and abs(a["area2d"] - b["area2d"]) <= 1e-12 * b["area2d"]
