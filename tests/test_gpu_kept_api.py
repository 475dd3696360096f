"""GPU: the reference's own entry points (include/delaunay_tri.h, include/gta_tri.h) and the g_tessla CLI,
served by libgtessla_b200.so, against the oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from g_tessla_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


class DTri(C.Structure):
    _fields_ = [("points", dp), ("npoints", C.c_int), ("triangles", ip), ("ntriangles", C.c_int), ("nverts", C.c_int)]


class TriArea(C.Structure):
    _fields_ = [("area", dp), ("area2D", dp), ("area2Dbox", dp), ("natoms", C.c_int), ("nframes", C.c_int)]


def test_dtriangulate_entry_point(gta, checker):
    L = gta.lib()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    L.dtinit()
    for n in (2, 3, 50, 1156, 6000):                        # 6000 > shared-memory capacity: large path
        p = np.ascontiguousarray(np.random.default_rng(n).random((n, 2)) * 9)
        t = DTri(p.ctypes.data_as(dp), n, None, -1, -1)
        L.dtriangulate(C.byref(t))
        want, _ = checker.triangulate(p)
        assert t.ntriangles == len(want) and t.nverts == n
        got = np.ctypeslib.as_array(t.triangles, shape=(t.ntriangles, 3)).copy() if t.ntriangles else np.zeros((0, 3), np.int32)
        libc.free(t.triangles)                                # caller frees, as with the reference (gta_tri.c:412)
        assert np.array_equal(got, want)
    # fewer than two points: message on stderr, struct untouched (reference delaunay_tri.c:416-421)
    p = np.zeros((1, 2))
    t = DTri(p.ctypes.data_as(dp), 1, None, -7, -9)
    L.dtriangulate(C.byref(t))
    assert t.ntriangles == -7 and t.nverts == -9 and not t.triangles
    assert L.dt_last_status() & 1


def test_delaunay_tessellate_entry_point(gta, checker):
    L = gta.lib()
    c = synth.config(2, nframes=50)
    F, N = c["xyz"].shape[:2]
    frames = [np.ascontiguousarray(c["xyz"][f]) for f in range(F)]          # one block per frame, like read_traj
    xptr = (dp * F)(*[fr.ctypes.data_as(dp) for fr in frames])
    box = np.zeros((F, 3, 3)); box[:, 0, 0] = c["box"][:, 0]; box[:, 1, 1] = c["box"][:, 1]; box[:, 2, 2] = c["box"][:, 2]
    a = TriArea(None, None, None, N, F)
    L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(TriArea), C.c_ubyte]
    L.delaunay_tessellate(xptr, box.ctypes.data, 0.8, 0, C.byref(a), 3)
    want = checker.tessellate(c["xyz"], c["box"], 0.8, 3)
    area = np.ctypeslib.as_array(a.area, shape=(F,)); a2 = np.ctypeslib.as_array(a.area2D, shape=(F,)); ab = np.ctypeslib.as_array(a.area2Dbox, shape=(F,))
    assert np.allclose(area, want["area"], rtol=1e-12, atol=0)
    assert np.allclose(a2, want["area2d"], rtol=1e-12, atol=0)
    assert np.array_equal(ab, want["areabox"])
    for f in range(F):
        assert np.array_equal(frames[f], c["xyz"][f])          # the caller's frames are not modified
    L.free_tri_area(C.byref(a))
    # one frame, no correction: delaunay_surface_area
    L.delaunay_surface_area.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_ubyte, dp, dp]
    a3, a2d = C.c_double(-1), C.c_double(-1)
    L.delaunay_surface_area(frames[0].ctypes.data, box[0].ctypes.data, N, 0, C.byref(a2d), C.byref(a3))
    w = checker.tessellate(c["xyz"][:1], c["box"][:1], 0.8, 2)
    assert abs(a3.value - w["area"][0]) <= 1e-12 * w["area"][0] and abs(a2d.value - w["area2d"][0]) <= 1e-12 * w["area2d"][0]


def _write_gro(path, xyz, box, fmt="%8.3f"):
    with open(path, "w") as f:
        for fr in range(len(xyz)):
            f.write("frame t= %d\n%5d\n" % (fr, xyz.shape[1]))
            for i, (x, y, z) in enumerate(xyz[fr]):
                f.write("%5d%-5s%5s%5d" % (i + 1, "POPC", "P", i + 1) + (fmt % x) + (fmt % y) + (fmt % z) + "\n")
            f.write("%10.5f%10.5f%10.5f\n" % tuple(box[fr]))


def _expected_dat(area, area2d, areabox, natoms):
    """The reference's print_areas format (src/gta_tri.c:448-473), restated."""
    out = []
    if area2d is not None:
        out.append('# FRAME\tAREA\t2DAREA\tBOX-AREA\t""/PARTICLE\n')
        for i in range(len(area)):
            out.append("%d\t%f\t%f\t%f\t%f\t%f\t%f\n" % (i, area[i], area2d[i], areabox[i], area[i] / natoms, area2d[i] / natoms, areabox[i] / natoms))
    else:
        out.append('# FRAME\tAREA\tBOX-AREA\t""/PARTICLE\n')
        for i in range(len(area)):
            out.append("%d\t%f\t%f\t%f\t%f\n" % (i, area[i], areabox[i], area[i] / natoms, areabox[i] / natoms))
    return "".join(out)


def test_cli_gro_ndx_dat(gta, checker, tmp_path):
    exe = os.path.join(ROOT, "g_tessla_b200", "g_tessla")
    assert os.path.exists(exe), "CLI not built (python -m g_tessla_b200.build)"
    c = synth.config(1, nframes=12)
    # trajectory with 80 atoms per frame of which an index group selects the 64 lipid P atoms; .gro keeps 3 decimals
    rng = np.random.default_rng(3)
    xyz = np.round(c["xyz"], 3); box = np.round(c["box"], 5)
    extra = np.round(rng.random((12, 16, 3)) * 6.0, 3)
    full = np.concatenate([extra[:, :8], xyz, extra[:, 8:]], axis=1)
    _write_gro(tmp_path / "traj.gro", full, box)
    with open(tmp_path / "index.ndx", "w") as f:
        f.write("[ P_atoms ]\n")
        ids = np.arange(9, 9 + 64)
        for k in range(0, 64, 15):
            f.write(" ".join(str(v) for v in ids[k:k + 15]) + "\n")
        f.write("[ other ]\n1 2 3\n")
    for flags, extra_args in ((1, ["-corr"]), (3, ["-corr", "-2d"]), (0, [])):
        r = subprocess.run([exe, "-f", "traj.gro", "-n", "index.ndx", "-o", "out.dat", "-espace", "0.8"] + extra_args,
                           cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        want = checker.tessellate(xyz, box, 0.8, flags)
        exp = _expected_dat(want["area"], want["area2d"], want["areabox"], 64)
        assert open(tmp_path / "out.dat").read() == exp
        assert "Average surface area: %f" % want["area"].mean() in r.stdout
        assert "Surface areas saved to out.dat" in r.stdout
    assert os.path.exists(tmp_path / "gta.log")


def test_readers_pdb_trr(gta, checker, tmp_path):
    import struct
    exe = os.path.join(ROOT, "g_tessla_b200", "g_tessla")
    c = synth.config(1, nframes=5)
    xyz = np.round(c["xyz"], 3); box = np.round(c["box"], 3)
    # multi-model PDB (Angstrom)
    with open(tmp_path / "t.pdb", "w") as f:
        for fr in range(5):
            f.write("CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1\nMODEL %8d\n" % (box[fr, 0] * 10, box[fr, 1] * 10, box[fr, 2] * 10, fr + 1))
            for i, (x, y, z) in enumerate(xyz[fr]):
                f.write("ATOM  %5d  P   POP A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, i + 1, x * 10, y * 10, z * 10))
            f.write("TER\nENDMDL\n")
    # double-precision TRR (XDR, big-endian)
    with open(tmp_path / "t.trr", "wb") as f:
        for fr in range(5):
            n = xyz.shape[1]
            f.write(struct.pack(">ii", 1993, 13) + struct.pack(">i", 12) + b"GMX_trn_file")
            f.write(struct.pack(">13i", 0, 0, 72, 0, 0, 0, 0, n * 24, 0, 0, n, fr, 0))
            f.write(struct.pack(">2d", float(fr), 0.0))
            b = np.zeros((3, 3)); b[0, 0], b[1, 1], b[2, 2] = c["box"][fr]
            f.write(b.astype(">f8").tobytes()); f.write(c["xyz"][fr].astype(">f8").tobytes())
    want_r = checker.tessellate(xyz, box, 0.8, 1)
    want_f = checker.tessellate(c["xyz"], c["box"], 0.8, 1)
    for name, want in (("t.pdb", want_r), ("t.trr", want_f)):
        r = subprocess.run([exe, "-f", name, "-o", name + ".dat", "-corr"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        assert open(tmp_path / (name + ".dat")).read() == _expected_dat(want["area"], None, want["areabox"], 64), name


def test_print_flag_writes_showme_files(gta, checker, tmp_path):
    """GTA_PRINT: triangles<k>.node / .ele per frame in the reference's format (src/gta_tri.c:355-360, :475-498)."""
    L = gta.lib()
    c = synth.config(1, nframes=3)
    F, N = c["xyz"].shape[:2]
    frames = [np.ascontiguousarray(c["xyz"][f]) for f in range(F)]
    xptr = (dp * F)(*[fr.ctypes.data_as(dp) for fr in frames])
    box = np.zeros((F, 3, 3)); box[:, 0, 0] = c["box"][:, 0]; box[:, 1, 1] = c["box"][:, 1]
    a = TriArea(None, None, None, N, F)
    L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(TriArea), C.c_ubyte]
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        L.delaunay_tessellate(xptr, box.ctypes.data, 0.8, 1, C.byref(a), 1 | 4)
    finally:
        os.chdir(cwd)
    want = checker.tessellate(c["xyz"], c["box"], 0.8, 1, want_corr=True)
    assert np.allclose(np.ctypeslib.as_array(a.area, shape=(F,)), want["area"], rtol=1e-12, atol=0)
    nodes = sorted(p for p in os.listdir(tmp_path) if p.endswith(".node"))
    assert len(nodes) == F
    k = int(nodes[0][len("triangles"):-len(".node")])
    p = np.concatenate([c["xyz"][0, :, :2], want["corr"][0, :int(want["ncorr"][0]), :2]])
    tri, _ = checker.triangulate(p)
    ele = open(tmp_path / ("triangles%d.ele" % k)).read().splitlines()
    assert ele[0] == "%d\t3\t0" % len(tri)
    assert ele[1:] == ["%d\t%d\t%d\t%d" % (i, t[0], t[1], t[2]) for i, t in enumerate(tri)]
    node = open(tmp_path / nodes[0]).read().splitlines()
    assert node[0] == "%d\t2\t0\t0" % len(p) and node[1] == "0\t%f\t%f" % (p[0, 0], p[0, 1])
    L.free_tri_area(C.byref(a))


def test_large_frames_with_correction_points(gta, checker):
    """Frames beyond the shared-memory kernels WITH -corr (a whole leaflet of 80 x 80 lipids: 6,400 atoms + 324 box-edge
    points; reference src/gta_tri.c:106-296 handles any size): the correction points are generated on the device by the
    batched kernels' own corr_points, and the frame goes through the large-frame path. gta_b200_tessellate_large and the
    kept delaunay_tessellate against the reference: areas within 1e-12, box area and point count equal; an atom outside the
    box is refused with the batched calls' status, not reproduced."""
    L = gta.lib()
    xyz, box = synth.jittered_bilayer(3, 80, seed=12)
    F, N = xyz.shape[:2]
    assert N + 4 + 4 * int(box[:, :2].max() / 0.8 + 1) > L.gta_b200_max_points()
    want = checker.tessellate(xyz, box, 0.8, 3, want_corr=True)
    box9 = np.zeros((F, 3, 3)); box9[:, 0, 0] = box[:, 0]; box9[:, 1, 1] = box[:, 1]; box9[:, 2, 2] = box[:, 2]
    for f in range(F):
        r = gta.tessellate_large(xyz[f], box9[f], 0.8, flags=3)
        assert r["status"] == 0 and r["ncorr"] == int(want["ncorr"][f]) and r["areabox"] == want["areabox"][f], (f, r)
        assert r["ntri"] == 2 * (N + r["ncorr"] - 1) - (r["ncorr"])          # the hull is the box: every -corr point is on it
        assert abs(r["area"] - want["area"][f]) <= 1e-12 * want["area"][f]
        assert abs(r["area2d"] - want["area2d"][f]) <= 1e-12 * want["area2d"][f]
    frames = [np.ascontiguousarray(xyz[f]) for f in range(F)]
    xptr = (dp * F)(*[fr.ctypes.data_as(dp) for fr in frames])
    a = TriArea(None, None, None, N, F)
    L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(TriArea), C.c_ubyte]
    L.delaunay_tessellate(xptr, box9.ctypes.data, 0.8, 0, C.byref(a), 3)
    area = np.ctypeslib.as_array(a.area, shape=(F,)); a2 = np.ctypeslib.as_array(a.area2D, shape=(F,)); ab = np.ctypeslib.as_array(a.area2Dbox, shape=(F,))
    assert np.allclose(area, want["area"], rtol=1e-12, atol=0) and np.allclose(a2, want["area2d"], rtol=1e-12, atol=0)
    assert np.array_equal(ab, want["areabox"])
    L.free_tri_area(C.byref(a))
    bad = xyz[0].copy(); bad[17, 0] = box[0, 0] * 1.5
    r = gta.tessellate_large(bad, box9[0], 0.8, flags=1)
    assert r["status"] & 16                                              # GTA_B200_ST_OUT_OF_BOX
    assert np.isnan(r["area"])


def test_cli_whole_system_frames_with_corr(gta, checker, tmp_path):
    """The command line on a trajectory whose frames are beyond the shared-memory kernels (no index group: 4,900 atoms per
    frame), with -corr -2d: tessellate_area -> delaunay_tessellate -> gta_b200_tessellate_large per frame; the .dat must be
    the file the reference's numbers give."""
    exe = os.path.join(ROOT, "g_tessla_b200", "g_tessla")
    xyz, box = synth.jittered_bilayer(2, 70, seed=21)
    xyz = np.round(xyz, 3); box = np.round(box, 5)
    xyz[:, :, 0] = np.clip(xyz[:, :, 0], 0.001, box[:, None, 0] - 0.001); xyz[:, :, 1] = np.clip(xyz[:, :, 1], 0.001, box[:, None, 1] - 0.001)
    assert xyz.shape[1] + 4 + 4 * int(box[:, :2].max() / 0.8 + 1) > gta.lib().gta_b200_max_points()
    _write_gro(tmp_path / "traj.gro", xyz, box)
    r = subprocess.run([exe, "-f", "traj.gro", "-o", "out.dat", "-espace", "0.8", "-corr", "-2d"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    want = checker.tessellate(xyz, box, 0.8, 3)
    assert open(tmp_path / "out.dat").read() == _expected_dat(want["area"], want["area2d"], want["areabox"], xyz.shape[1])


def test_large_frames_shard_over_the_gpus(gta, checker):
    """delaunay_tessellate on a trajectory of frames beyond the shared-memory kernels with every GPU of the box (one host
    thread per device, every ng-th frame): the same bits as on one GPU. Needs two devices (the driver's 1-GPU run skips it;
    scratch/large_traj.py is the measured form: 282 -> 482 frames/s of 10^5 atoms on two B200s)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    code = ("import ctypes as C, sys, numpy as np; sys.path.insert(0, %r); import g_tessla_b200 as g; from g_tessla_b200 import synth\n"
            "xyz, box = synth.jittered_bilayer(6, 70, seed=3); F, N = xyz.shape[:2]; L = g.lib(); dp = C.POINTER(C.c_double)\n"
            "class T(C.Structure): _fields_ = [('area', dp), ('area2D', dp), ('area2Dbox', dp), ('natoms', C.c_int), ('nframes', C.c_int)]\n"
            "fr = [np.ascontiguousarray(xyz[f]) for f in range(F)]; xp = (dp * F)(*[a.ctypes.data_as(dp) for a in fr])\n"
            "b9 = np.zeros((F, 3, 3)); b9[:, 0, 0] = box[:, 0]; b9[:, 1, 1] = box[:, 1]; b9[:, 2, 2] = box[:, 2]\n"
            "a = T(None, None, None, N, F); L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(T), C.c_ubyte]\n"
            "L.delaunay_tessellate(xp, b9.ctypes.data, 0.8, 0, C.byref(a), 1); print('AREAS', ' '.join(repr(float(v)) for v in np.ctypeslib.as_array(a.area, shape=(F,))))\n") % ROOT
    got = {}
    for ng in (1, 2):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GTA_B200_NGPUS=str(ng)), capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("AREAS")]
        assert r.returncode == 0 and line, r.stdout[-1000:] + r.stderr[-2000:]
        got[ng] = np.array([float(v) for v in line[0].split()[1:]])
    assert np.array_equal(got[1], got[2])
    xyz, box = synth.jittered_bilayer(6, 70, seed=3)
    want = checker.tessellate(xyz, box, 0.8, 1)
    assert np.allclose(got[2], want["area"], rtol=1e-12, atol=0)
