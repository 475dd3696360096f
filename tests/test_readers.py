"""CPU: the trajectory / index readers behind tessellate_area (include/gta_io.h) -- plain host code of the library,
callable without a GPU."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from xtc_writer import write_xtc

dp = C.POINTER(C.c_double)


def _read(path):
    import g_tessla_b200 as g
    L = g.lib()
    libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
    x = C.POINTER(dp)(); box = dp(); nf = C.c_int(0); na = C.c_int(0)
    L.read_traj.argtypes = [C.c_char_p, C.POINTER(C.POINTER(dp)), C.POINTER(dp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    L.read_traj(str(path).encode(), C.byref(x), C.byref(box), C.byref(nf), C.byref(na), None)
    F, N = nf.value, na.value
    xyz = np.array([np.ctypeslib.as_array(x[f], shape=(N, 3)).copy() for f in range(F)]) if F else np.zeros((0, 0, 3))
    b = np.ctypeslib.as_array(box, shape=(F, 3, 3)).copy() if F else np.zeros((0, 3, 3))
    for f in range(F):
        libc.free(x[f])
    if F:
        libc.free(x); libc.free(box)
    return xyz, b


@pytest.fixture()
def frames():
    rng = np.random.default_rng(12)
    xyz = np.round(rng.random((4, 50, 3)) * 6.0, 3)
    box = np.round(6.0 + rng.random((4, 3)), 3)
    return xyz, box


def test_gro_pdb_trr(frames, tmp_path):
    xyz, box = frames
    with open(tmp_path / "t.gro", "w") as f:
        for fr in range(4):
            f.write("t=%d\n%5d\n" % (fr, 50))
            for i, (x, y, z) in enumerate(xyz[fr]):
                f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (i + 1, "LIP", "P", i + 1, x, y, z))
            f.write("%10.5f%10.5f%10.5f\n" % tuple(box[fr]))
    with open(tmp_path / "t.pdb", "w") as f:
        for fr in range(4):
            f.write("CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1\nMODEL %8d\n" % (*(box[fr] * 10), fr + 1))
            for i, (x, y, z) in enumerate(xyz[fr]):
                f.write("ATOM  %5d  P   LIP A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, i + 1, x * 10, y * 10, z * 10))
            f.write("ENDMDL\n")
    for prec, fmt in ((8, ">f8"), (4, ">f4")):
        with open(tmp_path / ("t%d.trr" % prec), "wb") as f:
            for fr in range(4):
                f.write(struct.pack(">ii", 1993, 13) + struct.pack(">i", 12) + b"GMX_trn_file")
                f.write(struct.pack(">13i", 0, 0, 9 * prec, 0, 0, 0, 0, 50 * 3 * prec, 0, 0, 50, fr, 0))
                f.write(np.array([fr, 0.0], dtype=fmt).tobytes())
                b = np.zeros((3, 3)); b[0, 0], b[1, 1], b[2, 2] = box[fr]
                f.write(b.astype(fmt).tobytes()); f.write(xyz[fr].astype(fmt).tobytes())
    for name, tol in (("t.gro", 0), ("t.pdb", 1e-12), ("t8.trr", 0), ("t4.trr", None)):
        got, gb = _read(tmp_path / name)
        assert got.shape == xyz.shape, name
        want = xyz if tol is not None else xyz.astype(np.float32).astype(np.float64)
        wb = box if tol is not None else box.astype(np.float32).astype(np.float64)
        assert np.allclose(got, want, rtol=0, atol=tol or 0), name
        assert np.allclose(gb[:, [0, 1, 2], [0, 1, 2]], wb, rtol=0, atol=tol or 0), name


@pytest.mark.parametrize("runs", [False, True])
def test_xtc_round_trip(frames, tmp_path, runs):
    """The .xtc decoder against an independent encoder of the same published scheme (tests/xtc_writer.py): plain
    frames, and frames whose atoms come in runs of small offsets (clustered like water molecules)."""
    xyz, box = frames
    if runs:
        rng = np.random.default_rng(3)
        centers = rng.random((4, 10, 1, 3)) * 5.0 + 0.5
        xyz = np.round((centers + rng.normal(0, 0.05, (4, 10, 5, 3))).reshape(4, 50, 3), 3)
    want = write_xtc(tmp_path / "t.xtc", xyz, box, runs=runs, smallidx=30 if runs else None)
    if runs:                                                             # the run coding was really used: the file is smaller
        write_xtc(tmp_path / "plain.xtc", xyz, box, runs=False)
        assert os.path.getsize(tmp_path / "t.xtc") < 0.9 * os.path.getsize(tmp_path / "plain.xtc")
    got, gb = _read(tmp_path / "t.xtc")
    assert got.shape == xyz.shape
    assert np.array_equal(got, want)
    assert np.abs(got - xyz).max() <= 0.5e-3 + 1e-6                       # the format keeps 1/precision nm
    assert np.allclose(gb[:, [0, 1, 2], [0, 1, 2]], box.astype(np.float32))
    # tiny frames (<= 9 atoms) are stored as plain floats
    want = write_xtc(tmp_path / "s.xtc", xyz[:, :7], box)
    got, _ = _read(tmp_path / "s.xtc")
    assert np.array_equal(got, want)


def test_ndx_filter(tmp_path):
    import g_tessla_b200 as g
    L = g.lib()
    with open(tmp_path / "i.ndx", "w") as f:
        f.write("[ first ]\n   3   1\n 5\n[ second ]\n2 4\n")
    x = np.arange(2 * 6 * 3, dtype=np.float64).reshape(2, 6, 3)
    rows = [np.ascontiguousarray(x[f]) for f in range(2)]
    xp = (dp * 2)(*[r.ctypes.data_as(dp) for r in rows])
    newx = C.POINTER(dp)(); na = C.c_int(6)
    L.ndx_filter_traj.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.POINTER(dp)), C.c_int, C.POINTER(C.c_int)]
    L.ndx_filter_traj(str(tmp_path / "i.ndx").encode(), xp, C.byref(newx), 2, C.byref(na))
    assert na.value == 3
    for f in range(2):
        got = np.ctypeslib.as_array(newx[f], shape=(3, 3))
        assert np.array_equal(got, x[f][[2, 0, 4]])                      # first group only, 1-based, in index order


# ---- .xtc byte vectors built BY HAND from the published xdr3dfcoord layout (no encoder of this repo involved) -------------
# Frame header (XDR, big-endian): magic 1995, natoms, step 7, time 2.5f, box diag (6.5, 7.25, 8.0) as 9 floats; then
# natoms again, precision 1000.0f, minint[3], maxint[3], smallidx 9, byte count, the bit stream, zero padding to 4 bytes.
# Every atom is stored in full (no runs): its three offsets from minint as ONE mixed-radix number
# V = ((x - minx) * sy + (y - miny)) * sz + (z - minz) in bitsize = bit_length(sx * sy * sz) bits, least significant BYTE
# first (bits of a byte most significant first), followed by one flag bit 0 ("run length unchanged").
#   vector A: min (-3, 10, 100), max (3, 14, 108) -> sizes (7, 5, 9), 315 -> 9 bits + 1 flag bit per atom, 12 atoms = 15 bytes.
#     atom 0 = (-3, 10, 100): V = 0           -> 00000000 0 | 0              stream starts 0000000000
#     atom 1 = ( 3, 14, 108): V = 34*9+8 = 314 = 0x13a -> low byte 0x3a = 00111010, then the top bit 1 | flag 0
#     so the stream begins 00000000 00001110 1010....  = 00 0e a.: the vector below begins 00 0e a2.
#   vector B: min (-20000000, 0, 5), max (20000000, 3, 70000): a size above 2^24 - 1 switches to one field per coordinate
#     of bit_length(size) bits (26, 3, 17), most significant bit first, plus the flag bit: 47 bits per atom, 10 atoms = 59 bytes.
XTC_A = ("000007cb0000000c000000074020000040d0000000000000000000000000000040e80000000000000000000000000000410000000000000c"
         "447a0000fffffffd0000000a00000064000000030000000e0000006c000000090000000f000ea20830c813804208712e80c8ec00")
XTC_A_INTS = [(-3, 10, 100), (3, 14, 108), (3, 12, 100), (-3, 11, 103), (1, 12, 102), (-2, 13, 106), (-3, 10, 104), (-1, 14, 104),
              (-1, 12, 105), (1, 10, 106), (2, 14, 107), (-2, 11, 105)]
XTC_B = ("000007cb0000000a000000074020000040d0000000000000000000000000000040e80000000000000000000000000000410000000000000a"
         "447a0000feced300000000000000000501312d000000000300011170000000090000003b000000180001312d00088b582ef2427021416c42620a"
         "bb20693af5ad4bcb293db0406e8742d17681bb0250b20582984f4cdfc915681bb9de25a4c000")
XTC_B_INTS = [(-20000000, 3, 5), (20000000, 0, 70000), (-16923326, 3, 66073), (-8063951, 0, 21982), (-18275907, 3, 46388),
              (3406518, 0, 32994), (-12386281, 3, 33216), (-18786160, 1, 49489), (788095, 1, 17759), (-16365892, 2, 46237)]


@pytest.mark.parametrize("hexvec,ints", [(XTC_A, XTC_A_INTS), (XTC_B, XTC_B_INTS)])
def test_xtc_hand_built_vectors(hexvec, ints, tmp_path):
    """The .xtc decoder against byte vectors derived by hand from the format description (above): the decoded positions are
    exactly integer / precision in the format's single-precision arithmetic. (Still not a GROMACS-written file: DESIGN.md
    keeps the .xtc reader's parity "unpinned".)"""
    path = tmp_path / "hand.xtc"
    path.write_bytes(bytes.fromhex(hexvec) * 2)                 # the same frame twice: frame boundaries are found too
    xyz, box = _read(path)
    assert xyz.shape == (2, len(ints), 3)
    want = (np.array(ints, dtype=np.float32) * (np.float32(1.0) / np.float32(1000.0))).astype(np.float64)
    assert np.array_equal(xyz[0], want) and np.array_equal(xyz[1], want)
    assert np.array_equal(box[0].diagonal(), np.array([6.5, 7.25, 8.0]))
