"""GPU: parity at BASELINE's own frame counts, the throughput path un-forced, the multi-word exact stage on the device,
the reference's own print_areas, large frames behind the kept entry points, error reporting."""
import ctypes as C
import os
import subprocess
from fractions import Fraction as Fr

import numpy as np
import pytest

from g_tessla_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def _tri_sample(gta_result, checker, c, want, frames):
    for f in frames:
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]])
        tri, _ = checker.triangulate(p)
        assert np.array_equal(gta_result["tri"][f], tri), f          # the reference's triangles in its emission order


@pytest.mark.parametrize("cfg,nfr", [(2, 1000), (5, 10000)])
def test_baseline_frame_counts_vs_reference(gta, checker, cfg, nfr):
    """BASELINE.json's literal sizes: cfg 2 = 1,000 frames (-corr -2d), cfg 5 = 10,000 frames of the degenerate lattice.
    Every frame's areas against the CPU reference; triangles on a sample."""
    c = synth.config(cfg, nframes=nfr)
    want = checker.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], nthreads=0, want_corr=True)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["ncorr"], want["ncorr"])
    assert all(np.array_equal(r["corr"][f, :r["ncorr"][f]], want["corr"][f, :r["ncorr"][f]]) for f in range(nfr))
    assert np.array_equal(r["areabox"], want["areabox"])
    assert np.allclose(r["area"], want["area"], rtol=RTOL, atol=0)
    if cfg == 2:
        assert np.allclose(r["area2d"], want["area2d"], rtol=RTOL, atol=0)
        assert np.allclose(r["area2d"], r["areabox"], rtol=RTOL, atol=0)              # reference README.md:16
    sub = np.arange(0, nfr, max(1, nfr // 40))
    rs = gta.tessellate(c["xyz"][sub], c["box"][sub], c["espace"], c["flags"], want_tri=True)
    for k, f in enumerate(sub):
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][f, :int(want["ncorr"][f]), :2]])
        tri, _ = checker.triangulate(p)
        assert np.array_equal(rs["tri"][k], tri), (cfg, f)


@pytest.mark.parametrize("cfg,nfr", [(3, 6000), (5, 6000)])
def test_throughput_path_unforced_triangles(gta, checker, dc_path, cfg, nfr):
    """A batch above the default threshold takes the throughput path (pre-pass, walker kernel, area kernel) without any
    forcing, and hands back triangle lists from it: compared with the reference in emission order on a sample, areas
    against the fused kernel on every frame (the two paths must agree to the rounding of the summation tree)."""
    L = gta.lib()
    assert L.gta_b200_set_batch_threshold(-1) >= 0 and nfr >= 2048
    c = synth.config(cfg, nframes=nfr)

    def walker_launches():
        n = C.c_int(-1)
        assert L.gta_b200_stage_ms(-1, None, None, None, C.byref(n)) == 0
        return n.value
    L.gta_b200_stage_ms.argtypes = [C.c_int, dp, dp, dp, ip]
    L.gta_b200_stage_reset(-1)
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_tri=True, want_corr=True)
    assert (r["status"] == 0).all()
    assert walker_launches() >= 1                                       # the throughput path ran
    sub = np.arange(0, nfr, nfr // 30)
    want = checker.tessellate(c["xyz"][sub], c["box"][sub], c["espace"], c["flags"], nthreads=0, want_corr=True)
    for k, f in enumerate(sub):
        p = np.concatenate([c["xyz"][f, :, :2], want["corr"][k, :int(want["ncorr"][k]), :2]])
        tri, _ = checker.triangulate(p)
        assert np.array_equal(r["tri"][f], tri), (cfg, f)
    assert np.allclose(r["area"][sub], want["area"], rtol=RTOL, atol=0)
    old = L.gta_b200_set_batch_threshold(0)                              # the fused kernel on the same frames
    try:
        L.gta_b200_stage_reset(-1)
        q = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
        assert walker_launches() == 0
    finally:
        L.gta_b200_set_batch_threshold(old)
    assert np.allclose(r["area"], q["area"], rtol=RTOL, atol=0)
    assert np.array_equal(r["ntri"], q["ntri"])


def _orient_exact(v):
    ax, ay, bx, by, cx, cy = map(Fr, v.tolist())
    d = (ax - cx) * (by - cy) - (ay - cy) * (bx - cx)
    return (d > 0) - (d < 0)


def _incircle_exact(v):
    ax, ay, bx, by, cx, cy, dx, dy = map(Fr, v.tolist())
    adx, ady, bdx, bdy, cdx, cdy = ax - dx, ay - dy, bx - dx, by - dy, cx - dx, cy - dy
    d = ((adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) + (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy)
         + (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady))
    return (d > 0) - (d < 0)


def exact_stage_families():
    """Degenerate operands that neither the FP filter nor the FP exactness stage can settle, built to need 1, 2 and 4
    64-bit words in the integer stage: exactly collinear triples / exactly cocircular quadruples (rectangles) whose x are
    many-bit integers times 2^ex and whose y are many-bit integers times 2^ey -- every coordinate is exactly representable,
    the products are not, and ex - ey sets the bit spread. Half of each family is perturbed by one unit of the low scale."""
    rng = np.random.default_rng(17)
    fams = []
    for bits, ex, ey in ((29, 0, 0), (40, 20, -20), (40, 70, -70)):
        H, Lo = 2.0 ** ex, 2.0 ** ey
        o, ic = [], []
        for i in range(80):
            ax, ay, dx, dy = (float(v) for v in rng.integers(1 << (bits - 1), 1 << bits, 4))
            t = float(rng.integers(2, 9))
            cx, cy = ax + t * dx, ay + t * dy
            if i % 2: cy += 1.0
            o.append([ax * H, ay * Lo, (ax + dx) * H, (ay + dy) * Lo, cx * H, cy * Lo])
            x0, y0, w, h = (float(v) for v in rng.integers(1 << (bits - 1), 1 << bits, 4))
            q = [[x0, y0], [x0 + w, y0], [x0 + w, y0 + h], [x0, y0 + h]]
            if i % 2: q[3][1] += 1.0
            q = [q[k] for k in rng.permutation(4)]
            ic.append([v for pt in q for v in (pt[0] * H, pt[1] * Lo)])
        fams.append((np.array(o), np.array(ic)))
    return fams


def test_exact_stage_word_paths_on_device(gta, emu):
    """The integer exact stage picks 1, 2 or 4 64-bit words per coordinate from the operands' bit spread
    (gt_predicates.cuh). Each family is checked on the host build of the same header to really reach the integer stage
    with that word count, and must give the exact-rational sign ON THE GPU; operands spanning more than the 251
    representable bits must come back as the range error (2), never as a sign."""
    for word_idx, (o, ic) in enumerate(exact_stage_families()):
        emu.exact_stats()
        so_h = emu.signs("emu_orient2d_signs", o, 6); si_h = emu.signs("emu_incircle_signs", ic, 8)      # filter + A' + integers
        st = emu.exact_stats()
        assert st[word_idx] >= 30 and st[3 + word_idx] >= 40, (word_idx, st)       # the integer stage ran with this word count
        assert sum(st[:6]) == st[word_idx] + st[3 + word_idx], st
        want_o = np.array([_orient_exact(v) for v in o], dtype=np.int32)
        want_i = np.array([_incircle_exact(v) for v in ic], dtype=np.int32)
        assert (want_o == 0).sum() == 40 and (want_i == 0).sum() == 40              # the exact zeros are in there
        assert np.array_equal(so_h, want_o) and np.array_equal(si_h, want_i)
        assert np.array_equal(gta.api.orient2d_signs(o), want_o), word_idx          # the device
        assert np.array_equal(gta.api.incircle_signs(ic), want_i), word_idx
    # beyond 251 bits: the range error, and a frame with such a predicate is flagged PRECISION_RANGE (8)
    m = float((1 << 40) + 12345)
    far_o = np.array([[m * 2.0 ** 150, m * 2.0 ** -150, 2 * m * 2.0 ** 150, 2 * m * 2.0 ** -150, 3 * m * 2.0 ** 150, 3 * m * 2.0 ** -150]])
    assert emu.signs("emu_orient2d_signs", far_o, 6)[0] == 2
    assert gta.api.orient2d_signs(far_o)[0] == 2
    far_i = np.array([[m * 2.0 ** 150, m * 2.0 ** -150, 2 * m * 2.0 ** 150, m * 2.0 ** -150, 2 * m * 2.0 ** 150, 3 * m * 2.0 ** -150,
                       m * 2.0 ** 150, (3 * m + 1) * 2.0 ** -150]])
    assert gta.api.incircle_signs(far_i)[0] == 2
    # a frame whose predicates hit it: five exactly collinear points with that spread (distinct x: not duplicates)
    pts = np.array([[[k * m * 2.0 ** 150, k * m * 2.0 ** -150] for k in (1, 2, 3, 4, 5)]])
    _, st = gta.triangulate(pts)
    assert st[0] == 8, st


def test_cli_dat_equals_reference_print_areas(gta, ref, tmp_path):
    """The CLI's .dat against the file the reference's OWN print_areas (src/gta_tri.c:448-473, compiled into oracle/_ref)
    writes from the reference's own areas."""
    exe = os.path.join(ROOT, "g_tessla_b200", "g_tessla")
    c = synth.config(1, nframes=9)
    xyz = np.round(c["xyz"], 3); box = np.round(c["box"], 5)
    with open(tmp_path / "t.gro", "w") as f:
        for fr in range(len(xyz)):
            f.write("frame t= %d\n%5d\n" % (fr, xyz.shape[1]))
            for i, (x, y, z) in enumerate(xyz[fr]):
                f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (i + 1, "POPC", "P", i + 1, x, y, z))
            f.write("%10.5f%10.5f%10.5f\n" % tuple(box[fr]))
    for flags, args in ((1, ["-corr"]), (3, ["-corr", "-2d"])):
        r = subprocess.run([exe, "-f", "t.gro", "-o", "out.dat"] + args, cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        want = ref.tessellate(xyz, box, 0.8, flags)
        ref.print_areas(tmp_path / "ref.dat", want["area"], want["area2d"], want["areabox"], xyz.shape[1])
        assert open(tmp_path / "out.dat", "rb").read() == open(tmp_path / "ref.dat", "rb").read()


class TriArea(C.Structure):
    _fields_ = [("area", dp), ("area2D", dp), ("area2Dbox", dp), ("natoms", C.c_int), ("nframes", C.c_int)]


def test_large_frames_behind_kept_entry_points(gta, checker):
    """Frames beyond the shared-memory kernel (a whole system without -n) go through the large-frame path instead of
    coming back as zeros: delaunay_surface_area, and delaunay_tessellate without and with -corr (gta_b200_tessellate_large
    generates the box-edge points of such a frame), all against the reference."""
    L = gta.lib()
    n = L.gta_b200_max_points() + 1500
    rng = np.random.default_rng(9)
    xyz = np.ascontiguousarray(np.concatenate([rng.random((2, n, 2)) * 40.0, 4.0 + 0.2 * rng.standard_normal((2, n, 1))], axis=2))
    box = np.zeros((2, 3, 3)); box[:, 0, 0] = 40.0; box[:, 1, 1] = 40.0
    want = checker.tessellate(xyz, np.full((2, 3), 40.0), 0.8, 2)
    L.delaunay_surface_area.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_ubyte, dp, dp]
    a3, a2 = C.c_double(-1), C.c_double(-1)
    L.delaunay_surface_area(xyz[0].ctypes.data, box[0].ctypes.data, n, 0, C.byref(a2), C.byref(a3))
    assert L.gta_last_call_rc() == 0
    assert abs(a3.value - want["area"][0]) <= 1e-12 * want["area"][0] and abs(a2.value - want["area2d"][0]) <= 1e-12 * want["area2d"][0]
    frames = [np.ascontiguousarray(xyz[f]) for f in range(2)]
    xptr = (dp * 2)(*[fr.ctypes.data_as(dp) for fr in frames])
    L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(TriArea), C.c_ubyte]
    a = TriArea(None, None, None, n, 2)
    L.delaunay_tessellate(xptr, box.ctypes.data, 0.8, 0, C.byref(a), 2)
    assert L.gta_last_call_rc() == 0
    assert np.allclose(np.ctypeslib.as_array(a.area, shape=(2,)), want["area"], rtol=RTOL, atol=0)
    assert np.allclose(np.ctypeslib.as_array(a.area2D, shape=(2,)), want["area2d"], rtol=RTOL, atol=0)
    assert np.array_equal(np.ctypeslib.as_array(a.area2Dbox, shape=(2,)), np.array([1600.0, 1600.0]))
    L.free_tri_area(C.byref(a))
    b = TriArea(None, None, None, n, 2)
    L.delaunay_tessellate(xptr, box.ctypes.data, 0.8, 0, C.byref(b), 1)             # -corr on a frame this large: its own point stage
    assert L.gta_last_call_rc() == 0
    wc = checker.tessellate(xyz, np.full((2, 3), 40.0), 0.8, 1)
    assert np.allclose(np.ctypeslib.as_array(b.area, shape=(2,)), wc["area"], rtol=RTOL, atol=0)
    L.free_tri_area(C.byref(b))


def test_corr_point_count_cannot_overflow(gta):
    """A tiny espace asks for ~10^9 edge points per frame: a per-frame CAPACITY status, not an out-of-bounds access."""
    c = synth.config(1, nframes=6)
    L = gta.lib()
    F, N = c["xyz"].shape[:2]
    area = np.zeros(F); st = np.zeros(F, dtype=np.int32)
    import torch
    T = gta.Tessellator(F, N, N + 40, 1e-8, 1)
    T.run(torch.from_numpy(c["xyz"]).cuda(), torch.from_numpy(c["box"]).cuda(), 3)
    torch.cuda.synchronize()
    assert ((T.status.cpu().numpy() & 4) != 0).all() and torch.isnan(T.area).all()
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])                 # the device is fine afterwards
    assert (r["status"] == 0).all()


def test_stage_timers(gta, dc_path):
    """gta_b200_stage_ms: per-kernel device time of the throughput path since the last reset."""
    L = gta.lib()
    L.gta_b200_stage_ms.argtypes = [C.c_int, dp, dp, dp, ip]
    c = synth.config(3, nframes=4096)
    assert L.gta_b200_stage_reset(-1) == 0
    r = gta.tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    assert (r["status"] == 0).all()
    pre, walk, area, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    assert L.gta_b200_stage_ms(-1, C.byref(pre), C.byref(walk), C.byref(area), C.byref(n)) == 0
    assert n.value >= 1 and pre.value > 0 and area.value > 0 and walk.value > pre.value
    assert pre.value + walk.value + area.value <= r["kernel_ms"] * 1.05 + 0.5
