"""GPU: the CUDA path, called through the C-ABI, against the CPU oracle and the golden fixtures."""
import numpy as np
import pytest

from g_tessla_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12      # north_star: per-frame areas within 1e-12 relative in FP64


def _corr_equal(got, ncorr, want):
    """Appended -corr points bit-exact, frame by frame over each frame's own count."""
    return all(np.array_equal(got[f, :ncorr[f]], want[f, :ncorr[f]]) for f in range(len(ncorr)))


def _canon(t):
    return synth.canonical_triangles(t)


def test_predicate_signs(gta, golden):
    g = golden["predicates"]
    assert np.array_equal(gta.api.orient2d_signs(g["o2d"]), g["o2d_sign"])
    assert np.array_equal(gta.api.incircle_signs(g["icc"]), g["icc_sign"])


def test_golden_triangulations(gta, golden):
    g = golden["dtriangulate"]
    for k in g.files:
        if not k.startswith("pts_"):
            continue
        tri, st = gta.triangulate(g[k])
        assert st == 0, (k, st)
        want = g["tri_" + k[4:]]
        assert np.array_equal(_canon(tri), _canon(want)), k          # bit-exact index set
        assert np.array_equal(tri, want), k                          # and the reference's emission order


@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_golden_tessellate(gta, golden, cfg):
    t = golden["tessellate"]
    nfr = int(t["cfg%d_nframes" % cfg])
    c = synth.config(cfg, nframes=nfr)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_tri=True, want_corr=True)
    assert (r["status"] == 0).all(), r["status"]
    assert np.array_equal(r["ncorr"], t["cfg%d_ncorr" % cfg])
    assert _corr_equal(r["corr"], r["ncorr"], t["cfg%d_corr" % cfg])          # appended points bit-exact
    assert np.array_equal(r["areabox"], t["cfg%d_areabox" % cfg])
    assert np.allclose(r["area"], t["cfg%d_area" % cfg], rtol=RTOL, atol=0)
    if cfg == 2:
        assert np.allclose(r["area2d"], t["cfg2_area2d"], rtol=RTOL, atol=0)
        assert np.allclose(r["area2d"], r["areabox"], rtol=RTOL, atol=0)
    assert np.array_equal(_canon(r["tri"][0]), _canon(t["cfg%d_tri0" % cfg]))


@pytest.mark.parametrize("cfg,nfr", [(1, 1000), (2, 300), (3, 200), (5, 200)])
def test_configs_vs_oracle(gta, checker, cfg, nfr):
    c = synth.config(cfg, nframes=nfr)
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0, want_corr=True)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_tri=True, want_corr=True)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["ncorr"], want["ncorr"])
    assert _corr_equal(r["corr"], r["ncorr"], want["corr"])
    assert np.array_equal(r["areabox"], want["areabox"])
    assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    if cfg == 2:
        assert np.allclose(r["area2d"], want["area2d"], rtol=RTOL, atol=0)
    for f in range(0, nfr, max(1, nfr // 25)):
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]])
        tri, _ = checker.triangulate(p)
        assert np.array_equal(_canon(r["tri"][f]), _canon(tri)), (cfg, f)
        assert len(tri) == r["ntri"][f]


def test_raw_point_sets_many_sizes(gta, checker):
    rng = np.random.default_rng(21)
    for n in [2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33, 64, 100, 191, 192, 193, 511, 512, 513, 1156, 2047, 3000]:
        p = rng.random((3, n, 2)) * 10
        tris, st = gta.triangulate(p)
        assert (st == 0).all(), (n, st)
        for f in range(3):
            want, _ = checker.triangulate(p[f])
            assert np.array_equal(tris[f], want), (n, f)


def test_status_codes(gta):
    # duplicate point
    p = np.random.default_rng(1).random((2, 20, 2))
    p[1, 7] = p[1, 3]
    _, st = gta.triangulate(p)
    assert st[0] == 0 and st[1] & 2
    # non-finite
    p = np.random.default_rng(1).random((1, 20, 2)); p[0, 0, 0] = np.nan
    _, st = gta.triangulate(p)
    assert st[0] & 64
    # atom outside the box with -corr
    c = synth.config(1, nframes=2)
    c["xyz"][1, 5, 0] = c["box"][1, 0] * 3.0
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    assert r["status"][0] == 0 and r["status"][1] & 16 and np.isnan(r["area"][1])
    # a single point: the reference prints TRIANGULATION ERROR and produces nothing
    r = gta.tessellate(np.zeros((1, 1, 3)), None, flags=0)
    assert r["status"][0] & 1


def test_empty_and_chunked(gta, checker):
    r = gta.tessellate(np.zeros((0, 64, 3)), np.zeros((0, 3)))
    assert len(r["area"]) == 0
    c = synth.config(1, nframes=777)
    a = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nstreams=1)
    b = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nstreams=5)
    assert np.allclose(a["area"], b["area"], rtol=RTOL, atol=0)    # chunking never changes results (beyond the summation tree)
    # logical shards concatenated == unsharded (multi-GPU correctness on one GPU)
    parts = [gta.tessellate(c["xyz"][s], c["box"][s], c["espace"], c["flags"])["area"]
             for s in (slice(0, 300), slice(300, 777))]
    assert np.allclose(np.concatenate(parts), a["area"], rtol=RTOL, atol=0)


def test_full_size_properties(gta):
    """cfg 3 at a large frame count: size-independent properties (no oracle at this size)."""
    c = synth.config(3, nframes=4000)
    fl = c["flags"] | synth.GTA_2D
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], fl)
    assert (r["status"] == 0).all()
    assert np.allclose(r["area2d"], r["areabox"], rtol=RTOL, atol=0)      # reference README.md:16
    assert (r["area"] >= r["area2d"] * (1 - 1e-12)).all()                   # lifting never shrinks a triangle
    nx = (c["box"][:, 0] / 0.8).astype(int); ny = (c["box"][:, 1] / 0.8).astype(int)
    n = 1024 + 4 + 2 * nx + 2 * ny
    h = 4 + 2 * nx + 2 * ny                                                  # every -corr point is on the hull
    assert np.array_equal(r["ntri"], 2 * (n - 1) - h)                        # reference delaunay_tri.c:608
    # permuting the atoms of a frame permutes indices only: areas agree to rounding of the sum order
    perm = np.random.default_rng(0).permutation(1024)
    r2 = gta.tessellate(c["xyz"][:200, perm], c["box"][:200], c["espace"], fl)
    assert np.allclose(r2["area"], r["area"][:200], rtol=1e-9, atol=0)


def test_device_resident_entry(gta):
    import torch
    c = synth.config(3, nframes=64)
    want = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    T = gta.Tessellator(64, 1024, 1024 + 140, c["espace"], c["flags"])
    dx = torch.from_numpy(c["xyz"]).cuda(); db = torch.from_numpy(c["box"]).cuda()
    T.run(dx, db, 3)
    torch.cuda.synchronize()
    assert np.allclose(T.area.cpu().numpy(), want["area"], rtol=RTOL, atol=0)
    assert (T.status.cpu().numpy() == 0).all()


@pytest.fixture()
def lane_per_frame(gta):
    """Force every launch through the lane-per-frame pipeline (normally only batches >= 32768 frames)."""
    L = gta.lib()
    old = L.gta_b200_set_batch_threshold(1)
    oldp = gta.set_path(gta.PATH_DC)
    yield
    gta.set_path(oldp)
    L.gta_b200_set_batch_threshold(old)


@pytest.mark.parametrize("cfg,nfr", [(1, 600), (2, 300), (3, 150), (5, 150)])
def test_lane_per_frame_configs_vs_oracle(gta, checker, lane_per_frame, cfg, nfr):
    c = synth.config(cfg, nframes=nfr)
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0, want_corr=True)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True, want_tri=True)
    assert (r["status"] == 0).all()
    assert _corr_equal(r["corr"], r["ncorr"], want["corr"])
    assert np.array_equal(r["areabox"], want["areabox"])
    assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    if cfg == 2:
        assert np.allclose(r["area2d"], want["area2d"], rtol=RTOL, atol=0)
    for f in range(0, nfr, max(1, nfr // 30)):
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]])
        tri, _ = checker.triangulate(p)
        assert np.array_equal(r["tri"][f], tri), (cfg, f)            # the reference's triangles in its emission order


def test_lane_per_frame_raw_sets_and_status(gta, checker, lane_per_frame, golden):
    g = golden["dtriangulate"]
    for k in g.files:
        if not k.startswith("pts_") or len(g[k]) < 2:
            continue
        tris, st = gta.triangulate(np.stack([g[k]] * 3))
        assert (st == 0).all(), k
        for t in tris:
            assert np.array_equal(t, g["tri_" + k[4:]]), k
    rng = np.random.default_rng(33)
    for n in (2, 3, 4, 5, 9, 33, 200, 1156, 3000):
        p = rng.random((2, n, 2)) * 7
        tris, st = gta.triangulate(p)
        assert (st == 0).all(), n
        for f in range(2):
            assert np.array_equal(tris[f], checker.triangulate(p[f])[0]), n
    # rejected frames keep their status and NaN area, the others are unaffected
    c = synth.config(1, nframes=40)
    c["xyz"][7, 3] = c["xyz"][7, 9]
    c["xyz"][11, 5, 0] = -1.0
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    assert r["status"][7] & 2 and r["status"][11] & 16 and np.isnan(r["area"][7]) and np.isnan(r["area"][11])
    ok = np.ones(40, bool); ok[[7, 11]] = False
    want = checker.tessellate(c["xyz"][ok], c["box"][ok], c["espace"], c["flags"])
    assert np.allclose(r["area"][ok], want["area"], rtol=RTOL, atol=0)


def test_lane_per_frame_batch_larger_than_one_launch(gta, checker):
    """A batch above one launch's worth of slots (148 SMs x 576) is split into even launches by the device call and into
    even chunks by the host call: every frame must come out exactly as from a small batch (cfg 1 frames, tiled)."""
    import torch
    c = synth.config(1, nframes=500)
    reps = 240                                                  # 120,000 frames > 85,248
    xyz = np.tile(c["xyz"], (reps, 1, 1)); box = np.tile(c["box"], (reps, 1))
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0)
    n = c["xyz"].shape[1]
    maxp = n + 4 + 4 * int(c["box"][:, :2].max() / c["espace"] + 1)
    T = gta.Tessellator(len(xyz), n, maxp, c["espace"], c["flags"])
    T.run(torch.from_numpy(xyz).cuda(), torch.from_numpy(box).cuda(), 3)
    torch.cuda.synchronize()
    assert int((T.status != 0).sum()) == 0
    area = T.area.cpu().numpy().reshape(reps, -1)
    assert (area == area[0]).all()                              # same frame, same bits, whichever launch it went into
    assert np.allclose(area[0], want["area"], rtol=RTOL, atol=0)
    assert np.array_equal(T.areabox.cpu().numpy()[:500], want["areabox"])
    assert (T.ntri.cpu().numpy().reshape(reps, -1) == T.ntri.cpu().numpy()[:500]).all()
    gta.star_counters(reset=True)
    r = gta.tessellate(xyz, box, c["espace"], c["flags"])       # host buffers: chunked pipeline
    assert (r["status"] == 0).all()
    d = r["area"].reshape(reps, -1) != area
    assert not d.any(), (int(d.sum()), gta.star_counters(), np.unique(np.nonzero(d)[0])[:8], np.unique(np.nonzero(d)[1])[:8])


def test_lane_per_frame_two_host_threads_one_device(gta, checker, lane_per_frame):
    """Two host threads, two streams, one device: the lane-per-frame launches share the device's workspace and must
    queue behind each other, not run side by side in it."""
    import threading
    import torch
    cs = [synth.config(1, nframes=400), synth.config(3, nframes=60)]
    wants = [checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0)["area"] for c in cs]
    got = [None, None]
    errs = []

    def work(k):
        try:
            c = cs[k]
            n = c["xyz"].shape[1]
            maxp = n + 4 + 4 * int(c["box"][:, :2].max() / c["espace"] + 1)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                xyz = torch.from_numpy(c["xyz"]).cuda(); box = torch.from_numpy(c["box"]).cuda()
                T = gta.Tessellator(len(c["xyz"]), n, maxp, c["espace"], c["flags"])
                for _ in range(6):
                    T.run(xyz, box, 3)
                stream.synchronize()
                assert int((T.status != 0).sum()) == 0
                got[k] = T.area.cpu().numpy()
        except Exception as e:                                  # surfaced in the main thread
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k in range(2):
        assert np.allclose(got[k], wants[k], rtol=RTOL, atol=0), k


def test_lane_per_frame_both_builds_agree(gta, checker, lane_per_frame):
    """The lane-per-frame kernel is built with and without the predicates' FP exactness stage and the host picks from
    the previous launch's count of undecided predicates: a degenerate batch (cfg 5) switches the next launch to the
    other build, a generic one (cfg 3) switches back. Every launch must give the reference's result, bit for bit the
    same from both builds."""
    c5 = synth.config(5, nframes=120); c3 = synth.config(3, nframes=120)
    w5 = checker.tessellate(c5["xyz"], c5["box"], c5["espace"], c5["flags"], nthreads=0)["area"]
    w3 = checker.tessellate(c3["xyz"], c3["box"], c3["espace"], c3["flags"], nthreads=0)["area"]
    runs5, runs3 = [], []
    for _ in range(3):                                          # 5, 5, 3, 3, 5, 5, ...: each build sees each kind of input
        for c, runs in ((c5, runs5), (c5, runs5), (c3, runs3), (c3, runs3)):
            r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
            assert (r["status"] == 0).all()
            runs.append(r["area"])
    assert all(np.array_equal(a, runs5[0]) for a in runs5) and all(np.array_equal(a, runs3[0]) for a in runs3)
    assert np.allclose(runs5[0], w5, rtol=RTOL, atol=0) and np.allclose(runs3[0], w3, rtol=RTOL, atol=0)
