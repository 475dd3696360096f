import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import g_tessla_b200 as g
        return g.lib().gta_b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip: only skip when not selected
    pass


@pytest.fixture(scope="session")
def port():
    from oracle import PortOracle
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref); prebuilt .so travels to the GPU box."""
    from oracle import RefOracle, ref_available, build_oracles
    if not ref_available():
        if os.path.isdir("/root/reference/src"):
            build_oracles(ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return RefOracle()


@pytest.fixture(scope="session")
def checker(port):
    """Best available CPU checker: the compiled reference if present, else the pinned port."""
    from oracle import RefOracle, ref_available
    return RefOracle() if ref_available() else port


@pytest.fixture(scope="session")
def emu():
    """Host build of the device headers (tests/emu/emu_dt.cpp) -- test infrastructure."""
    d = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(d, "libemu_dt.so")
    srcs = [os.path.join(d, "emu_dt.cpp")] + [os.path.join(ROOT, "g_tessla_b200", "csrc", f)
                                               for f in ("dt_core.cuh", "gt_predicates.cuh", "lpf_core.cuh", "star_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               "-o", so, srcs[0]])
    L = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.emu_triangulate.argtypes = [dp, C.c_int, ip]
    L.emu_triangulate_lpf2.argtypes = [dp, C.c_int, C.c_int, C.c_uint, ip, C.POINTER(C.c_long)]
    L.emu_exact_stats.argtypes = [C.POINTER(C.c_long)]
    L.emu_star.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, ip, C.POINTER(C.c_long)]
    L.emu_star_tiles.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, ip, C.POINTER(C.c_long)]
    for n in ("emu_orient2d_signs", "emu_incircle_signs", "emu_orient2d_exact", "emu_incircle_exact"):
        getattr(L, n).argtypes = [dp, C.c_int, ip]

    class Emu:
        def triangulate(self, p):
            p = np.ascontiguousarray(p, dtype=np.float64)
            out = np.empty(3 * 2 * max(len(p), 1), dtype=np.int32)
            nt = L.emu_triangulate(p.ctypes.data_as(dp), len(p), out.ctypes.data_as(ip))
            assert nt >= 0, "emu error %d" % nt
            return out[:3 * nt].reshape(-1, 3).copy()

        def triangulate_walkers(self, p, level, seed):
            """The walker state machine of the throughput kernel: 2^level subtree walkers stepped one iteration at a time
            in a pseudo-random interleaving (seed != 0). Returns (triangles, [iterations, pool edges, longest, root])."""
            p = np.ascontiguousarray(p, dtype=np.float64)
            out = np.empty(3 * 2 * max(len(p), 1), dtype=np.int32)
            st = (C.c_long * 4)()
            nt = L.emu_triangulate_lpf2(p.ctypes.data_as(dp), len(p), int(level), int(seed), out.ctypes.data_as(ip), st)
            assert nt >= 0, "emu error %d" % nt
            return out[:3 * nt].reshape(-1, 3).copy(), list(st)

        def star(self, p, w0=2, dens=1.0, full=True, ccap=64):
            """The star path (star_core.cuh): (triangles in any order or None when the frame is not certified, status, counters)."""
            p = np.ascontiguousarray(p, dtype=np.float64)
            out = np.empty(3 * 2 * max(len(p), 1), dtype=np.int32)
            st = (C.c_long * 17)()
            nt = L.emu_star(p.ctypes.data_as(dp), len(p), int(w0), float(dens), 1 if full else 0, int(ccap), out.ctypes.data_as(ip), st)
            if nt < 0:
                return None, -nt, list(st)
            return out[:3 * nt].reshape(-1, 3).copy(), 0, list(st)

        def star_tiles(self, p, tw=32, halo=6, dens=1.0):
            """The star path over staged tiles of the grid (ls_tile_kernel's logic): (triangles or None, status, [deferred, tiles, hull])."""
            p = np.ascontiguousarray(p, dtype=np.float64)
            out = np.empty(3 * 2 * max(len(p), 1), dtype=np.int32)
            st = (C.c_long * 4)()
            nt = L.emu_star_tiles(p.ctypes.data_as(dp), len(p), int(tw), int(halo), float(dens), out.ctypes.data_as(ip), st)
            if nt < 0:
                return None, -nt, list(st)
            return out[:3 * nt].reshape(-1, 3).copy(), 0, list(st)

        def exact_stats(self):
            """[orient2d 1/2/4 words, incircle 1/2/4 words, orient2d total, incircle total] since the last call."""
            st = (C.c_long * 8)()
            L.emu_exact_stats(st)
            return list(st)

        def signs(self, fn, p, w):
            p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, w)
            out = np.empty(len(p), dtype=np.int32)
            getattr(L, fn)(p.ctypes.data_as(dp), len(p), out.ctypes.data_as(ip))
            return out
    return Emu()


@pytest.fixture(scope="session")
def golden():
    return {k: np.load(os.path.join(GOLDEN, k + ".npz")) for k in ("dtriangulate", "tessellate", "predicates")}


@pytest.fixture(scope="session")
def gta():
    """The product binding; on the GPU box the library must load and see a device."""
    import g_tessla_b200 as g
    L = g.lib()
    assert L.gta_b200_device_count() > 0, "no CUDA device visible to libgtessla_b200.so"
    return g


@pytest.fixture()
def dc_path(gta):
    """Batched calls through the divide-and-conquer engine only (tests of that engine's own kernels)."""
    old = gta.set_path(gta.PATH_DC)
    yield
    gta.set_path(old)


@pytest.fixture()
def star_path(gta):
    """Batched calls through the star engine even when triangle lists are asked for (they then come in any order)."""
    old = gta.set_path(gta.PATH_STAR)
    yield
    gta.set_path(old)
