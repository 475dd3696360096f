"""ctypes binding of libgtessla_b200.so (include/gta_b200.h).

Mirrors the reference's call shapes for the hot path: `tessellate` = delaunay_tessellate
(reference src/gta_tri.c:80) over a whole trajectory, `triangulate` = dtriangulate
(reference src/delaunay_tri.c:413) over a batch of raw 2-D point sets.  Everything runs in
the CUDA library; if it is missing or no device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

GTA_CORRECT = 1   # reference include/gta_tri.h:24-28
GTA_2D = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None

STATUS_NAMES = {1: "too_few_points", 2: "duplicate", 4: "capacity", 8: "precision_range", 16: "out_of_box",
                32: "pool", 64: "nonfinite", 128: "near_tie_x", 256: "internal"}


class GtaError(RuntimeError):
    pass


def lib_path() -> str:
    # (GTA_B200_LIB: an experiment build of the same sources, see scratch/README.md)
    return os.environ.get("GTA_B200_LIB") or os.path.join(_HERE, "libgtessla_b200.so")


def lib() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise GtaError(path + " is missing: build it with `python -m g_tessla_b200.build` "
                       "(nvcc, sm_100a). There is no CPU implementation of this path.")
    L = C.CDLL(path)
    L.gta_b200_last_error.restype = C.c_char_p
    L.gta_b200_last_kernel_ms.restype = C.c_double
    L.gta_b200_max_points_for.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint]
    L.gta_b200_launch_shape.argtypes = [C.c_int, C.c_int, _ip, C.POINTER(C.c_size_t), _ip]
    L.gta_b200_tessellate_device.argtypes = [
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gta_b200_tessellate_batch.argtypes = [
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
        C.c_void_p, C.c_int, C.c_void_p]
    L.gta_b200_tessellate_multi.argtypes = [
        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gta_b200_triangulate_large.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, _ip, _ip]
    L.gta_b200_large_last_ms.restype = C.c_double
    L.gta_b200_fp64_peak_tflops.restype = C.c_double
    L.gta_b200_fp64_peak_tflops.argtypes = [C.c_int]
    L.gta_b200_stage_ms.argtypes = [C.c_int, _dp, _dp, _dp, _ip]
    L.gta_b200_stage_reset.argtypes = [C.c_int]
    L.gta_b200_stage_ms4.argtypes = [C.c_int, _dp, _ip]
    L.gta_b200_last_fallback_frames.argtypes = [C.c_int]
    L.gta_b200_last_fallback_frames.restype = C.c_longlong
    L.gta_b200_set_path.argtypes = [C.c_int]
    L.gta_b200_star_counters.argtypes = [C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_int]
    L.gta_b200_tessellate_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gta_b200_surface_area_large.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_int, _dp, _dp, _ip, _ip]
    L.gta_b200_tessellate_large.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_uint, C.c_int, _dp, _dp, _dp, _ip, _ip, _ip]
    L.gta_b200_orient2d_signs.argtypes = [_dp, C.c_int, _ip, C.c_int]
    L.gta_b200_incircle_signs.argtypes = [_dp, C.c_int, _ip, C.c_int]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise GtaError("libgtessla_b200 error %d: %s" % (rc, lib().gta_b200_last_error().decode()))


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def tessellate(xyz, box, espace=0.8, flags=GTA_CORRECT, device=-1, nstreams=0, ngpus=1,
               want_tri=False, want_corr=False):
    """delaunay_tessellate over host arrays. xyz [F,N,3] (or [F,N,2]), box [F,3] / [F,9] / [F,2].

    Returns dict(area, area2d, areabox, status[, tri, ntri, corr, ncorr]); tri is a list of
    [ntri,3] int32 arrays in the reference's emission order."""
    L = lib()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    F, N, D = xyz.shape
    bstride = 0
    if box is not None:
        box = np.ascontiguousarray(box, dtype=np.float64)
        bstride = int(np.prod(box.shape[1:]))
        box = box.reshape(F, bstride)
        if bstride not in (2, 3, 9):
            raise ValueError("box must be [F,2], [F,3] (diagonal) or [F,9] (matrix)")
    area = np.zeros(F); areabox = np.zeros(F)
    area2d = np.zeros(F) if (flags & GTA_2D) else None
    status = np.zeros(F, dtype=np.int32)
    out = dict(area=area, area2d=area2d, areabox=areabox, status=status)
    if ngpus > 1:
        if want_tri or want_corr:
            raise ValueError("triangle / correction-point output is single-device")
        _check(L.gta_b200_tessellate_multi(_vp(xyz), D, _vp(box), bstride, F, N, float(espace), int(flags),
                                           int(ngpus), int(nstreams), _vp(area), _vp(area2d), _vp(areabox), _vp(status)))
        return out
    tri = ntri = corr = ncorr = None
    tri_cap = corr_cap = 0
    maxp = L.gta_b200_max_points_for(None if box is None else box.ctypes.data_as(_dp), bstride, F, N, float(espace), int(flags))
    if want_tri:
        tri_cap = max(2 * maxp, 1)
        tri = np.zeros((F, tri_cap, 3), dtype=np.int32)
    ntri = np.zeros(F, dtype=np.int32)
    if want_corr and (flags & GTA_CORRECT):
        corr_cap = max(maxp - N, 1)
        corr = np.zeros((F, corr_cap, 3)); ncorr = np.zeros(F, dtype=np.int32)
    _check(L.gta_b200_tessellate_batch(_vp(xyz), D, _vp(box), bstride, F, N, float(espace), int(flags),
                                       int(device), int(nstreams), _vp(area), _vp(area2d), _vp(areabox), _vp(status),
                                       _vp(tri), tri_cap, _vp(ntri), _vp(corr), corr_cap, _vp(ncorr)))
    out["ntri"] = ntri
    if want_tri:
        out["tri"] = [tri[f, :ntri[f]].copy() for f in range(F)]
    if corr is not None:
        out["corr"] = corr; out["ncorr"] = ncorr
    out["kernel_ms"] = L.gta_b200_last_kernel_ms()
    out["launches"] = L.gta_b200_last_launches()
    return out


def triangulate(points, device=-1):
    """dtriangulate over a batch of equal-sized raw 2-D point sets [F,n,2] (or one set [n,2]).

    Returns (list of [ntri,3] int32 triangle arrays in reference emission order, status[F])."""
    p = np.ascontiguousarray(points, dtype=np.float64)
    single = p.ndim == 2
    if single:
        p = p[None]
    r = tessellate(p, None, flags=0, device=device, want_tri=True)
    return (r["tri"][0], int(r["status"][0])) if single else (r["tri"], r["status"])


def triangulate_large(points, device=-1):
    """dtriangulate on one large raw 2-D point set [n,2] (the global-memory path, any n >= 2).

    Returns (triangles [ntri,3] int32 in reference emission order, status, device_ms)."""
    p = np.ascontiguousarray(points, dtype=np.float64)
    n = len(p)
    tri = np.zeros((max(2 * n, 1), 3), dtype=np.int32)
    nt = C.c_int(0); st = C.c_int(0)
    _check(lib().gta_b200_triangulate_large(_vp(p), n, int(device), _vp(tri), len(tri), C.byref(nt), C.byref(st)))
    return tri[:nt.value].copy(), st.value, lib().gta_b200_large_last_ms()


def surface_area_large(xyz, flags=0, device=-1):
    """delaunay_surface_area on one large frame of atoms [n,3] without correction points (any n >= 2).

    Returns dict(area, area2d, ntri, status, device_ms, engine): engine 1 = star engine, 0 = level-by-level divide and conquer."""
    p = np.ascontiguousarray(xyz, dtype=np.float64)
    a3, a2, nt, st = C.c_double(0.0), C.c_double(0.0), C.c_int(0), C.c_int(0)
    _check(lib().gta_b200_surface_area_large(_vp(p), len(p), int(flags), int(device), C.byref(a3), C.byref(a2), C.byref(nt), C.byref(st)))
    return dict(area=a3.value, area2d=a2.value, ntri=nt.value, status=st.value, device_ms=lib().gta_b200_large_last_ms(),
                engine=lib().gta_b200_large_last_engine())


def tessellate_large(xyz, box, espace, flags=0, device=-1):
    """One large frame the way delaunay_tessellate treats a frame: -corr points generated on the device (flags & CORRECT),
    then the large-frame path. xyz [n,3], box [3,3] or [9]. Returns dict(area, area2d, areabox, ntri, ncorr, status, engine)."""
    p = np.ascontiguousarray(xyz, dtype=np.float64)
    b = np.ascontiguousarray(box, dtype=np.float64).reshape(-1)
    a3, a2, ab, nt, nc, st = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_int(0), C.c_int(0), C.c_int(0)
    _check(lib().gta_b200_tessellate_large(_vp(p), len(p), b.ctypes.data_as(_dp), float(espace), int(flags), int(device),
                                           C.byref(a3), C.byref(a2), C.byref(ab), C.byref(nt), C.byref(nc), C.byref(st)))
    return dict(area=a3.value, area2d=a2.value, areabox=ab.value, ntri=nt.value, ncorr=nc.value, status=st.value,
                engine=lib().gta_b200_large_last_engine())


def orient2d_signs(pts, device=-1):
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 6)
    out = np.zeros(len(p), dtype=np.int32)
    _check(lib().gta_b200_orient2d_signs(p.ctypes.data_as(_dp), len(p), out.ctypes.data_as(_ip), device))
    return out


def incircle_signs(pts, device=-1):
    p = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 8)
    out = np.zeros(len(p), dtype=np.int32)
    _check(lib().gta_b200_incircle_signs(p.ctypes.data_as(_dp), len(p), out.ctypes.data_as(_ip), device))
    return out


def stage_ms(device=-1, reset=False):
    """Per-kernel device milliseconds of the batched path since the last reset: dict(star, prepass, walk, area, launches)."""
    L = lib()
    if reset:
        _check(L.gta_b200_stage_reset(int(device)))
        return None
    t, n = (C.c_double * 4)(), C.c_int()
    _check(L.gta_b200_stage_ms4(int(device), t, C.byref(n)))
    return dict(star=t[0], prepass=t[1], walk=t[2], area=t[3], launches=n.value)


PATH_AUTO, PATH_DC, PATH_STAR = 0, 1, 2


def set_path(mode):
    """Engine of the batched calls (include/gta_b200.h): PATH_AUTO, PATH_DC (divide and conquer only), PATH_STAR (stars also
    when triangle lists are asked for). Returns the previous mode."""
    return lib().gta_b200_set_path(int(mode))


def star_counters(device=-1, reset=False):
    """(frames certified by their stars, frames handed to the divide-and-conquer engine) on `device` since the last reset."""
    a, b = C.c_ulonglong(), C.c_ulonglong()
    _check(lib().gta_b200_star_counters(int(device), C.byref(a), C.byref(b), 1 if reset else 0))
    return a.value, b.value


def last_fallback_frames(device=-1):
    """Frames the most recent sampled star launch handed to the divide-and-conquer engine (-1: no star launch yet)."""
    return int(lib().gta_b200_last_fallback_frames(int(device)))


class Tessellator:
    """Device-resident arm (inputs already in HBM): torch owns the memory and the stream, the
    library enqueues its kernel on that stream (gta_b200_tessellate_device)."""

    def __init__(self, nframes, natoms, max_points, espace=0.8, flags=GTA_CORRECT, device="cuda:0", in_dim=3):
        import torch
        self.torch = torch
        self.dev = torch.device(device)
        self.F, self.N, self.maxp, self.espace, self.flags, self.in_dim = nframes, natoms, max_points, espace, flags, in_dim
        kw = dict(device=self.dev)
        self.area = torch.zeros(nframes, dtype=torch.float64, **kw)
        self.area2d = torch.zeros(nframes, dtype=torch.float64, **kw)
        self.areabox = torch.zeros(nframes, dtype=torch.float64, **kw)
        self.status = torch.zeros(nframes, dtype=torch.int32, **kw)
        self.ntri = torch.zeros(nframes, dtype=torch.int32, **kw)
        self.work = torch.zeros(4, dtype=torch.int32, **kw)

    def run(self, d_xyz, d_box, box_stride):
        """Enqueue one pass over [F,N,in_dim] device coordinates on torch's current stream."""
        torch = self.torch
        st = torch.cuda.current_stream(self.dev).cuda_stream
        with torch.cuda.device(self.dev):
            _check(lib().gta_b200_tessellate_device(
                d_xyz.data_ptr(), self.in_dim, 0 if d_box is None else d_box.data_ptr(), box_stride,
                self.F, self.N, self.maxp, float(self.espace), int(self.flags),
                self.area.data_ptr(), self.area2d.data_ptr() if (self.flags & GTA_2D) else None,
                self.areabox.data_ptr(), self.status.data_ptr(), None, 0, self.ntri.data_ptr(),
                None, 0, None, self.work.data_ptr(), st))
