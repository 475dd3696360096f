"""GPU: the star engine (csrc/gta_star.cuh, star_core.cuh) -- the default of the batched calls -- against the CPU reference:
triangle sets bit-exact after canonical sorting, areas within 1e-12 relative, and the hand-over of frames it does not
certify (cocircular ties, duplicates, refused frames) to the divide-and-conquer engine inside the same call."""
import numpy as np
import pytest

from g_tessla_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _canon(t):
    return synth.canonical_triangles(t)


@pytest.mark.parametrize("cfg,nfr", [(1, 1000), (2, 400), (3, 300)])
def test_star_triangles_and_areas_vs_reference(gta, checker, star_path, cfg, nfr):
    c = synth.config(cfg, nframes=nfr)
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0, want_corr=True)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True, want_tri=True)
    assert (r["status"] == 0).all()
    assert gta.last_fallback_frames() == 0                              # general position: every frame certified by its stars
    assert np.array_equal(r["ncorr"], want["ncorr"])
    assert all(np.array_equal(r["corr"][f, :r["ncorr"][f]], want["corr"][f, :r["ncorr"][f]]) for f in range(nfr))
    assert np.array_equal(r["areabox"], want["areabox"])
    assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    if cfg == 2:
        assert np.allclose(r["area2d"], want["area2d"], rtol=RTOL, atol=0)
        assert np.allclose(r["area2d"], r["areabox"], rtol=RTOL, atol=0)          # reference README.md:16
    for f in range(0, nfr, max(1, nfr // 40)):
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]])
        tri, _ = checker.triangulate(p)
        assert r["ntri"][f] == len(tri)
        assert np.array_equal(_canon(r["tri"][f]), _canon(tri)), (cfg, f)


def test_star_is_the_default_and_agrees_with_divide_and_conquer(gta):
    """No triangle list asked for: the default engine is the stars (one launch + three empty fallback launches per chunk);
    the divide-and-conquer engine on the same frames gives the same counts and areas to the rounding of the sum order."""
    c = synth.config(3, nframes=3000)
    fl = c["flags"] | synth.GTA_2D
    gta.stage_ms(reset=True)
    a = gta.tessellate(c["xyz"], c["box"], c["espace"], fl)
    t = gta.stage_ms()
    assert (a["status"] == 0).all() and t["star"] > 0 and t["star"] > 5 * t["walk"]
    assert gta.last_fallback_frames() == 0
    old = gta.set_path(gta.PATH_DC)
    try:
        b = gta.tessellate(c["xyz"], c["box"], c["espace"], fl)
    finally:
        gta.set_path(old)
    assert np.array_equal(a["ntri"], b["ntri"]) and np.array_equal(a["areabox"], b["areabox"])
    assert np.allclose(a["area"], b["area"], rtol=RTOL, atol=0) and np.allclose(a["area2d"], b["area2d"], rtol=RTOL, atol=0)
    assert np.max(np.abs(a["area"] - b["area"]) / b["area"]) < 1e-14


def test_tied_frames_fall_back_inside_the_call(gta, checker):
    """cfg 5 (lattice: cocircular ties in every frame) mixed into cfg-3-like frames of the same size: the tied frames are
    listed by the star kernel and finished by pre-pass + walkers + area kernel, each frame equal to the reference."""
    c5 = synth.config(5, nframes=40)
    n = c5["xyz"].shape[1]
    rng = np.random.default_rng(4)
    gen = c5["xyz"].copy()
    gen[:, :, :2] += (rng.random((40, n, 2)) - 0.5).astype(np.float32).astype(np.float64) * 0.05      # ties broken: general position
    gen[:, :, :2] = np.clip(gen[:, :, :2], 1e-3, c5["box"][:, None, :2] - 1e-3)
    xyz = np.empty((80, n, 3)); box = np.empty((80, 3)); tied = np.zeros(80, bool)
    xyz[0::2] = c5["xyz"]; xyz[1::2] = gen; box[0::2] = c5["box"]; box[1::2] = c5["box"]; tied[0::2] = True      # interleave tied / generic
    want = checker.tessellate(xyz, box, c5["espace"], c5["flags"], nthreads=0, want_corr=True)
    r = gta.tessellate(xyz, box, c5["espace"], c5["flags"], want_corr=True)
    assert (r["status"] == 0).all()
    fb = gta.last_fallback_frames()
    assert tied.sum() <= fb <= tied.sum() + 3                         # (a jittered frame may still hold a tie by accident)
    assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    assert np.array_equal(r["areabox"], want["areabox"]) and np.array_equal(r["ncorr"], want["ncorr"])
    nt = [len(checker.triangulate(np.concatenate([xyz[f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]]))[0]) for f in range(0, 80, 7)]
    assert np.array_equal(r["ntri"][::7], nt)


def test_all_tied_batches_skip_the_stars_after_a_while(gta, checker):
    """A trajectory whose every frame is tied (cfg 5) costs the stars for nothing: after a launch that handed more than half
    of its frames over, the next launches go straight to the divide-and-conquer engine. Results do not change."""
    c = synth.config(5, nframes=600)
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0)
    for _ in range(3):
        r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
        assert (r["status"] == 0).all()
        assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    gta.stage_ms(reset=True)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    t = gta.stage_ms()
    assert t["star"] < 0.05 * t["walk"]                                # no star launch any more
    # ... and a generic trajectory gets its stars back (the skip counts launches down)
    g = synth.config(1, nframes=64)
    for _ in range(20):
        gta.tessellate(g["xyz"], g["box"], g["espace"], g["flags"])
    gta.stage_ms(reset=True)
    gta.tessellate(g["xyz"], g["box"], g["espace"], g["flags"])
    assert gta.stage_ms()["star"] > 0


def test_star_raw_point_sets_without_corr(gta, checker, star_path):
    """Raw 2-D point sets (no -corr: irregular hulls, hull vertices scan the whole grid), sizes 2 ... 3000."""
    rng = np.random.default_rng(12)
    for n in (2, 3, 4, 5, 9, 33, 200, 1156, 3000):
        p = rng.random((3, n, 2)) * np.array([5.0, 2.0])
        tris, st = gta.triangulate(p)
        assert (st == 0).all(), n
        for f in range(3):
            want, _ = checker.triangulate(p[f])
            assert np.array_equal(_canon(tris[f]), _canon(want)), (n, f)


def test_star_statuses_come_from_the_fallback(gta, star_path):
    p = np.random.default_rng(1).random((3, 20, 2))
    p[1, 7] = p[1, 3]                                                   # duplicate point
    p[2, 0, 0] = np.nan
    _, st = gta.triangulate(p)
    assert st[0] == 0 and st[1] & 2 and st[2] & 64
    c = synth.config(1, nframes=3)
    c["xyz"][1, 5, 0] = c["box"][1, 0] * 3.0                            # atom outside the box with -corr
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    assert r["status"][0] == 0 and r["status"][2] == 0 and r["status"][1] & 16 and np.isnan(r["area"][1])


def test_star_golden_sets(gta, star_path, golden):
    """The reference's own triangulations of the golden point sets (near-degenerate ones included): frames the stars certify
    must hold exactly those triangles; the others come from the fallback in the reference's emission order."""
    g = golden["dtriangulate"]
    for k in g.files:
        if not k.startswith("pts_") or len(g[k]) < 2:
            continue
        tris, st = gta.triangulate(np.stack([g[k]] * 2))
        assert (st == 0).all(), k
        for t in tris:
            assert np.array_equal(_canon(t), _canon(g["tri_" + k[4:]])), k


def test_star_device_resident_large_batch_properties(gta):
    """cfg 3 at 20,000 device-resident frames: every frame certified, projected area == box area, triangle count rule."""
    import torch
    c = synth.config(3, nframes=2000)
    reps = 10
    xyz = torch.from_numpy(np.tile(c["xyz"], (reps, 1, 1))).cuda(); box = torch.from_numpy(np.tile(c["box"], (reps, 1))).cuda()
    fl = c["flags"] | synth.GTA_2D
    T = gta.Tessellator(len(xyz), 1024, 1024 + 140, c["espace"], fl)
    T.run(xyz, box, 3)
    torch.cuda.synchronize()
    assert (T.status == 0).all() and gta.last_fallback_frames() == 0
    a2, ab = T.area2d.cpu().numpy(), T.areabox.cpu().numpy()
    assert np.allclose(a2, ab, rtol=RTOL, atol=0)
    nx = (c["box"][:, 0] / 0.8).astype(int); ny = (c["box"][:, 1] / 0.8).astype(int)
    h = 4 + 2 * nx + 2 * ny
    assert np.array_equal(T.ntri.cpu().numpy(), np.tile(2 * (1024 + h - 1) - h, reps))
    a = T.area.cpu().numpy()
    assert np.array_equal(a[:2000], a[2000:4000])                      # same frame, same CTA-independent result
