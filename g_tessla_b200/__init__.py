"""g_tessla_b200 -- B200-native drop-in for g_tessla's per-frame Delaunay + area hot path.

The product is the C-ABI library `libgtessla_b200.so` (include/gta_b200.h, include/gta_tri.h,
include/delaunay_tri.h) built from g_tessla_b200/csrc/ for sm_100a.  This package is the thin
Python binding used by tests/ and bench.py; it never imports oracle/ and has no CPU path.
"""
from .api import (  # noqa: F401
    GTA_CORRECT, GTA_2D, PATH_AUTO, PATH_DC, PATH_STAR, Tessellator, lib, lib_path, tessellate, triangulate, triangulate_large, GtaError,
    set_path, last_fallback_frames, stage_ms, star_counters, surface_area_large, tessellate_large,
)
