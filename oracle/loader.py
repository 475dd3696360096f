"""ctypes bindings for the CPU oracles. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libgtessla_ref.so")
REF_WRAP_SO = os.path.join(HERE, "_ref", "libgtessla_ref_wrap.so")
PORT_SO = os.path.join(HERE, "liboracle_port.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def build_oracles(ref: bool | None = None) -> None:
    """Compile the port always and oracle/_ref when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class _TessOracle:
    """Shared call shapes for the two oracle libraries (prefix 'ref_' or 'port_')."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        f = getattr(L, prefix + "tessellate")
        f.restype = C.c_double
        f.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint,
                      _dp, _dp, _dp, _dp, C.c_int, _ip]
        self._tess = f
        g = getattr(L, prefix + "triangulate")
        g.restype = C.c_int
        g.argtypes = [_dp, C.c_int, _ip, _ip]
        self._tri = g
        for name, n in (("orient2d_batch", 6), ("incircle_batch", 8)):
            h = getattr(L, prefix + name)
            h.restype = None
            h.argtypes = [_dp, C.c_int, _dp]
        mt = getattr(L, prefix + "max_threads")
        mt.restype = C.c_int
        self.max_threads = mt

    def triangulate(self, points: np.ndarray):
        """points [n,2] f64 -> (triangles [ntri,3] int32 in reference emission order, nverts)."""
        pts = np.ascontiguousarray(points, dtype=np.float64)
        n = pts.shape[0]
        out = np.empty(3 * max(2 * (n - 1), 1), dtype=np.int32)
        nv = C.c_int(-1)
        nt = self._tri(_p(pts), n, _p(out, _ip), C.byref(nv))
        if nt < 0:
            return None, nv.value
        return out[:3 * nt].reshape(-1, 3).copy(), nv.value

    def tessellate(self, xyz, box, espace, flags, nthreads=1, want_corr=False):
        """xyz [F,N,3], box [F,3] -> dict(area, area2d, areabox, corr, ncorr, seconds)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        box = np.ascontiguousarray(box, dtype=np.float64)
        F, N, _ = xyz.shape
        area = np.zeros(F); areabox = np.zeros(F)
        area2d = np.zeros(F) if (flags & 2) else None
        corr = ncorr = None
        cap = 0
        if want_corr and (flags & 1):
            cap = int(4 + 2 * (box[:, 0].max() // espace + 1) + 2 * (box[:, 1].max() // espace + 1))
            corr = np.zeros((F, cap, 3)); ncorr = np.zeros(F, dtype=np.int32)
        sec = self._tess(_p(xyz), _p(box), F, N, float(espace), int(nthreads), int(flags),
                         _p(area), _p(area2d), _p(areabox), _p(corr), cap, _p(ncorr, _ip))
        return dict(area=area, area2d=area2d, areabox=areabox, corr=corr, ncorr=ncorr, seconds=sec)

    def orient2d(self, p: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 6)
        out = np.empty(len(p))
        getattr(self.lib, self.prefix + "orient2d_batch")(_p(p), len(p), _p(out))
        return out

    def incircle(self, p: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 8)
        out = np.empty(len(p))
        getattr(self.lib, self.prefix + "incircle_batch")(_p(p), len(p), _p(out))
        return out


class RefOracle(_TessOracle):
    """The unmodified reference (oracle/_ref). wrap=True loads the counter build."""

    def __init__(self, wrap: bool = False):
        path = REF_WRAP_SO if wrap else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        super().__init__(path, "ref_")
        self.lib.ref_counters.argtypes = [C.POINTER(C.c_longlong), C.c_int]
        if wrap:
            self.lib.ref_set_dump.argtypes = [C.c_char_p]

    def print_areas(self, fname, area, area2d, areabox, natoms):
        """The reference's own print_areas (src/gta_tri.c:448-473) -> fname."""
        area = np.ascontiguousarray(area, dtype=np.float64); areabox = np.ascontiguousarray(areabox, dtype=np.float64)
        a2 = None if area2d is None else np.ascontiguousarray(area2d, dtype=np.float64)
        f = self.lib.ref_print_areas
        f.restype = None
        f.argtypes = [C.c_char_p, _dp, _dp, _dp, C.c_int, C.c_int]
        f(str(fname).encode(), _p(area), _p(a2), _p(areabox), len(area), int(natoms))

    def counters(self, reset=True):
        out = (C.c_longlong * 4)()
        self.lib.ref_counters(out, int(reset))
        return dict(orient2d=out[0], orient2d_zero=out[1], incircle=out[2], incircle_zero=out[3])

    def set_dump(self, path):
        self.lib.ref_set_dump(None if path is None else path.encode())


class PortOracle(_TessOracle):
    """This repo's own C restatement (oracle/oracle_port.c)."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build_oracles(ref=False)
        super().__init__(PORT_SO, "port_")
