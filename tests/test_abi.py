"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"static\s+inline[^{;]*\{.*?\n\}", "", txt, flags=re.S)   # inline helpers are not exported
        for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", txt, flags=re.M):
            syms.add(m.group(1))
    return syms


def test_library_exports_declared_symbols():
    import g_tessla_b200 as g
    L = g.lib()
    syms = _declared_symbols()
    assert {"gta_b200_tessellate_batch", "gta_b200_tessellate_device", "gta_b200_tessellate_multi",
            "gta_b200_tessellate_frames"} <= syms
    missing = [s for s in sorted(syms) if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_path_without_device():
    """Without a CUDA device every compute entry point must fail loudly (no fallback)."""
    import g_tessla_b200 as g
    if g.lib().gta_b200_device_count() > 0:
        pytest.skip("device present")
    with pytest.raises(g.GtaError, match="no CUDA device"):
        g.tessellate(np.zeros((1, 4, 3)), np.ones((1, 3)))
    with pytest.raises(g.GtaError, match="no CUDA device"):
        g.api.orient2d_signs(np.zeros((1, 6)))


def test_product_does_not_use_oracle():
    """Nothing under g_tessla_b200/ may import, link or dlopen anything under oracle/."""
    pat = re.compile(r"^\s*(import oracle|from oracle|#include\s+\"[^\"]*oracle)|liboracle|libgtessla_ref", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "g_tessla_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), (dirpath, f)


def test_max_points_for_mirrors_reference_formula():
    import g_tessla_b200 as g
    L = g.lib()
    box = np.array([[6.4, 6.41, 8.0], [25.6, 25.9, 8.0]])
    dp = C.POINTER(C.c_double)
    assert L.gta_b200_max_points_for(box.ctypes.data_as(dp), 3, 2, 64, 0.8, 1) == 64 + 4 + 2 * 32 + 2 * 32
    assert L.gta_b200_max_points_for(box.ctypes.data_as(dp), 3, 1, 64, 0.8, 1) == 64 + 4 + 2 * 8 + 2 * 8
    assert L.gta_b200_max_points_for(box.ctypes.data_as(dp), 3, 2, 64, 0.8, 0) == 64
