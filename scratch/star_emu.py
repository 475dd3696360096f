"""Star path (star_core.cuh, host build) against the oracle: parity on the configs + work counters.  python scratch/star_emu.py [w0] [dens]"""
import sys, os, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import conftest
from g_tessla_b200 import synth
from oracle import RefOracle, PortOracle, ref_available
emu = conftest.emu.__wrapped__()
chk = RefOracle() if ref_available() else PortOracle()
w0 = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dens = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
full = (sys.argv[3] != '0') if len(sys.argv) > 3 else True
names = "cand left icc exact expand steps points hullcert nhull gx gy".split()
for cfg, nfr in ((1, 20), (3, 20), (5, 4)):
    c = synth.config(cfg, nframes=nfr)
    r = chk.tessellate(c["xyz"], c["box"], c["espace"], c["flags"], want_corr=True)
    tot = np.zeros(17); bad = 0; cert = 0
    for f in range(nfr):
        nc = int(r["ncorr"][f])
        p = np.concatenate([c["xyz"][f, :, :2], r["corr"][f, :nc, :2]])
        want, _ = chk.triangulate(p)
        tri, st, cnt = emu.star(p, w0, dens, full)
        tot += np.array(cnt)
        if tri is None:
            continue
        cert += 1
        if not np.array_equal(synth.canonical_triangles(tri), synth.canonical_triangles(want)):
            bad += 1
    n = len(p)
    print("cfg %d n=%d: certified %d/%d, mismatches %d" % (cfg, n, cert, nfr, bad))
    print("   per point:", {k: round(v / tot[6], 2) for k, v in zip(names[:8], tot[:8])}, "grid", int(tot[9] / nfr), int(tot[10] / nfr))
