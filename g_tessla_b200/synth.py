"""Seeded synthetic bilayer trajectories for the BASELINE.json configs.

All generators use numpy's counter-based Philox stream keyed by the config number so
that tests, bench.py and the golden-vector script see identical bytes.  Coordinates are
generated in FP64 and then rounded to the float32 grid (they stay stored as FP64, the
`real = double` contract of this repo), atoms lie strictly inside (0,bx) x (0,by), and
every frame is checked against the reference's input contract: no two points closer
than 1e-12 in both x and y (reference src/delaunay_tri.c:440-452 is broken for
duplicates) and no 0 < |dx| < 1e-12 (reference compareVerts, src/delaunay_tri.c:115-123).
"""
from __future__ import annotations

import numpy as np

GTA_CORRECT = 1  # reference include/gta_tri.h:24-28
GTA_2D = 2
GTA_PRINT = 4


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=int(seed)))


def _f32(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float32).astype(np.float64)


def jittered_bilayer(nframes: int, grid: int, seed: int, cell: float = 0.8,
                     jitter: float = 0.35, box_sigma: float = 0.005,
                     z0: float = 4.0, z_sigma: float = 0.15):
    """One leaflet of a bilayer: grid x grid lipids on a jittered lattice.

    Returns (xyz [nframes, grid*grid, 3] f64, box [nframes, 3] f64)."""
    rng = _rng(seed)
    natoms = grid * grid
    scale = 1.0 + box_sigma * rng.standard_normal((nframes, 2))
    box = np.empty((nframes, 3), dtype=np.float64)
    box[:, 0:2] = _f32(grid * cell * scale)
    box[:, 2] = _f32(np.full(nframes, 2.0 * z0))
    iy, ix = np.divmod(np.arange(natoms), grid)
    jit = rng.uniform(-jitter, jitter, size=(nframes, natoms, 2))
    fx = (ix[None, :] + 0.5 + jit[:, :, 0]) / grid
    fy = (iy[None, :] + 0.5 + jit[:, :, 1]) / grid
    xyz = np.empty((nframes, natoms, 3), dtype=np.float64)
    xyz[:, :, 0] = _f32(fx * box[:, None, 0])
    xyz[:, :, 1] = _f32(fy * box[:, None, 1])
    xyz[:, :, 2] = _f32(z0 + z_sigma * rng.standard_normal((nframes, natoms)))
    # keep atoms strictly inside the box after rounding
    tiny = np.float64(np.finfo(np.float32).tiny)
    for d in (0, 1):
        hi = np.nextafter(box[:, None, d].astype(np.float32), np.float32(0)).astype(np.float64)
        xyz[:, :, d] = np.clip(xyz[:, :, d], tiny * 1e10, hi)
    return xyz, box


def gel_lattice(nframes: int, seed: int, nx: int = 32, ny: int = 32, a: float = 0.48,
                defect_frac: float = 0.06, vacancy_frac: float = 0.02):
    """cfg 5: gel-phase near-hexagonal lattice with exact degeneracies.

    All coordinates are multiples of 2^-10 nm (exactly representable, products exact in
    FP64), rows are offset by a/2, a fraction of sites carries a small dyadic jitter,
    square 'defect domains' replace the triangular packing locally (exact cocircular
    quadruples) and some sites are vacant (moved onto a dyadic off-lattice position so
    natoms stays constant).  Returns (xyz, box)."""
    rng = _rng(seed)
    q = 2.0 ** -10
    aq = round(a / q) * q                      # lattice constant snapped
    rowq = round(a * np.sqrt(3.0) / 2.0 / q) * q   # row pitch snapped
    natoms = nx * ny
    bx = nx * aq
    by = ny * rowq
    j, i = np.divmod(np.arange(natoms), nx)
    base_x = (i + 0.25 + 0.5 * (j % 2)) * aq
    base_y = (j + 0.5) * rowq
    xyz = np.empty((nframes, natoms, 3), dtype=np.float64)
    box = np.empty((nframes, 3), dtype=np.float64)
    for f in range(nframes):
        x = base_x.copy()
        y = base_y.copy()
        # square defect domain: un-shear a block of rows so sites form exact squares/rectangles
        for _ in range(int(rng.integers(2, 5))):
            bw = int(rng.integers(3, 9)); bh = int(rng.integers(3, 9))
            i0 = int(rng.integers(1, nx - bw - 1)); j0 = int(rng.integers(1, ny - bh - 1))
            blk = (i >= i0) & (i < i0 + bw) & (j >= j0) & (j < j0 + bh)
            x[blk] = (i[blk] + 0.25) * aq
        # dyadic jitter on a subset
        m = rng.random(natoms) < defect_frac
        x[m] += rng.integers(-24, 25, size=int(m.sum())) * q
        y[m] += rng.integers(-24, 25, size=int(m.sum())) * q
        # vacancies: move the site by a dyadic off-lattice offset (keeps natoms fixed)
        v = rng.random(natoms) < vacancy_frac
        x[v] += (rng.integers(0, 2, size=int(v.sum())) * 2 - 1) * 0.375 * aq
        y[v] += (rng.integers(0, 2, size=int(v.sum())) * 2 - 1) * 0.3125 * rowq
        x = np.round(np.clip(x, q, bx - q) / q) * q
        y = np.round(np.clip(y, q, by - q) / q) * q
        xyz[f, :, 0] = x
        xyz[f, :, 1] = y
        xyz[f, :, 2] = 4.0 + np.round(0.1 * rng.standard_normal(natoms) / q) * q
        box[f] = (bx, by, 8.0)
        # resolve exact duplicates by nudging on the dyadic grid
        for _ in range(8):
            key = np.round(xyz[f, :, 0] / q).astype(np.int64) * (1 << 20) + np.round(xyz[f, :, 1] / q).astype(np.int64)
            _, first, counts = np.unique(key, return_index=True, return_counts=True)
            if (counts == 1).all():
                break
            dup = np.ones(natoms, dtype=bool); dup[first] = False
            xyz[f, dup, 0] = np.clip(xyz[f, dup, 0] + q * rng.integers(1, 9, size=int(dup.sum())), q, bx - q)
    return xyz, box


def uniform_patch(npoints: int, seed: int, side: float = 800.0, float_rounded: bool = True):
    """cfg 4: raw 2-D point set, uniform in a side x side patch. Returns [npoints, 2] f64."""
    rng = _rng(seed)
    pts = rng.uniform(0.0, side, size=(npoints, 2))
    if float_rounded:
        pts = _f32(pts)
    return pts


def check_input_contract(points_xy: np.ndarray) -> None:
    """Raise if a frame violates the reference's input contract (see module docstring)."""
    p = np.asarray(points_xy, dtype=np.float64)
    order = np.lexsort((p[:, 1], p[:, 0]))
    s = p[order]
    dx = np.diff(s[:, 0])
    dy = np.diff(s[:, 1])
    if np.any((np.abs(dx) < 1e-12) & (np.abs(dy) < 1e-12)):
        raise ValueError("duplicate points (within 1e-12) violate the reference input contract")
    if np.any((dx > 0) & (dx < 1e-12)):
        raise ValueError("0 < |dx| < 1e-12 makes the reference comparator non-transitive")


def config(cfg: int, nframes: int | None = None):
    """Return dict(xyz, box, espace, flags, name) for BASELINE.json config number 1,2,3,5."""
    if cfg in (1, 2):
        xyz, box = jittered_bilayer(nframes or 1000, 8, seed=1)
        flags = GTA_CORRECT | (GTA_2D if cfg == 2 else 0)
        name = "cfg%d: 128-lipid bilayer leaflet (64 P atoms), -corr -espace 0.8%s" % (cfg, " -2d" if cfg == 2 else "")
    elif cfg == 3:
        xyz, box = jittered_bilayer(nframes or 100000, 32, seed=3)
        flags = GTA_CORRECT
        name = "cfg3: 2048-lipid bilayer leaflet (1024 P atoms), -corr -espace 0.8"
    elif cfg == 5:
        xyz, box = gel_lattice(nframes or 10000, seed=5)
        flags = GTA_CORRECT
        name = "cfg5: gel-phase lattice with cocircular degeneracies (1024 sites), -corr"
    else:
        raise ValueError("cfg must be 1, 2, 3 or 5 (cfg 4 is uniform_patch)")
    return dict(xyz=xyz, box=box, espace=0.8, flags=flags, name=name)


def canonical_triangles(tri: np.ndarray) -> np.ndarray:
    """Sort the 3 indices of each triangle, then sort rows lexicographically."""
    t = np.sort(np.asarray(tri, dtype=np.int32).reshape(-1, 3), axis=1)
    if len(t) == 0:
        return t
    order = np.lexsort((t[:, 2], t[:, 1], t[:, 0]))
    return t[order]
