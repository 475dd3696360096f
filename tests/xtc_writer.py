"""A small .xtc ENCODER for the reader tests (test infrastructure). Written from the published description of the
xdr3dfcoord scheme, independently of the C++ decoder in g_tessla_b200/csrc/gta_host.cu. It produces valid streams of two
kinds: every atom with the frame-wide bit width (run = 0), and runs of small offsets (with the water swap of the first
two atoms of a run) at a fixed small index."""
import struct

import numpy as np

MAGICINTS = [0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
             1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
             65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122,
             1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216]
FIRSTIDX = 9


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, nbits, value):
        for i in range(nbits - 1, -1, -1):           # most significant bit first
            self.bits.append((value >> i) & 1)

    def put_ints(self, nbits, sizes, nums):
        v = (nums[0] * sizes[1] + nums[1]) * sizes[2] + nums[2]      # mixed radix, first digit most significant
        while nbits > 8:                              # least significant byte first, then the remaining high bits
            self.put(8, v & 0xff); v >>= 8; nbits -= 8
        if nbits > 0:
            self.put(nbits, v & ((1 << nbits) - 1))

    def bytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(int("".join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))


def sizeofints(sizes):
    prod = sizes[0] * sizes[1] * sizes[2]
    # bits of the mixed-radix maximum, counted the way the format does (whole bytes + bits of the top byte)
    nbytes = (max(prod.bit_length(), 1) + 7) // 8
    top = prod >> (8 * (nbytes - 1))
    nb = 0
    num = 1
    while top >= num:
        nb += 1; num *= 2
    return nb + (nbytes - 1) * 8


def write_xtc(path, xyz, box, precision=1000.0, runs=False, smallidx=None):
    """xyz [F,N,3] (nm), box [F,3] diagonal. Returns the coordinates a decoder must produce (float32 grid)."""
    out = np.zeros_like(np.asarray(xyz, dtype=np.float64))
    with open(path, "wb") as f:
        for fr in range(len(xyz)):
            n = xyz.shape[1]
            f.write(struct.pack(">iii", 1995, n, fr) + struct.pack(">f", float(fr)))
            b = np.zeros((3, 3), dtype=">f4"); b[0, 0], b[1, 1], b[2, 2] = box[fr]
            f.write(b.tobytes())
            f.write(struct.pack(">i", n))
            if n <= 9:
                f.write(np.asarray(xyz[fr], dtype=">f4").tobytes())
                out[fr] = np.asarray(xyz[fr], dtype=np.float32)
                continue
            ints = np.rint(np.asarray(xyz[fr], dtype=np.float64) * precision).astype(np.int64)
            mn, mx = ints.min(0), ints.max(0)
            sizes = [int(mx[c] - mn[c] + 1) for c in range(3)]
            assert max(sizes) <= 0xffffff
            bitsize = sizeofints(sizes)
            si = smallidx if smallidx is not None else 40
            small = MAGICINTS[si]; smallnum = small // 2
            w = BitWriter()
            i = 0
            first = True
            while i < n:
                # how many atoms after this one can follow as small offsets (chain: a1 big; a0 rel a1; a2 rel a0; a3 rel a2 ...)
                group = [i]
                if runs and i + 1 < n:
                    cand = [i + 1, i]                  # stream order: big = atom i+1, first small = atom i
                    ok = np.all(np.abs(ints[i] - ints[i + 1]) < smallnum)
                    if ok:
                        chain_prev = i
                        j = i + 2
                        while j < n and len(cand) < 9 and np.all(np.abs(ints[j] - ints[chain_prev]) < smallnum):
                            cand.append(j); chain_prev = j; j += 1
                        group = cand
                if len(group) == 1:
                    w.put_ints(bitsize, sizes, [int(v) for v in ints[i] - mn])
                    if first or run_state != 0:
                        w.put(1, 1); w.put(5, 0 + 1)          # run = 0, small index unchanged  (value % 3 == 1)
                    else:
                        w.put(1, 0)
                    run_state = 0; first = False
                    i += 1
                else:
                    big = group[0]
                    w.put_ints(bitsize, sizes, [int(v) for v in ints[big] - mn])
                    nsmall = len(group) - 1
                    if first or run_state != 3 * nsmall:
                        w.put(1, 1); w.put(5, 3 * nsmall + 1)
                    else:
                        w.put(1, 0)
                    run_state = 3 * nsmall; first = False
                    prev = ints[big]
                    for k, a in enumerate(group[1:]):
                        w.put_ints(si, [small] * 3, [int(v) for v in ints[a] - prev + smallnum])
                        prev = ints[a]
                    i += len(group)
            payload = w.bytes()
            f.write(struct.pack(">f", precision))
            f.write(struct.pack(">iii", *[int(v) for v in mn]) + struct.pack(">iii", *[int(v) for v in mx]))
            f.write(struct.pack(">i", si))
            f.write(struct.pack(">i", len(payload)) + payload + b"\0" * (-len(payload) % 4))
            inv = np.float32(1.0) / np.float32(precision)
            out[fr] = (ints.astype(np.float32) * inv).astype(np.float64)
    return out
