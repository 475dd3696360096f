"""One large frame WITH -corr (a leaflet of G x G lipids) through gta_b200_tessellate_large, against the reference.  python scratch/large_corr.py [G]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
from oracle import RefOracle, PortOracle, ref_available
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
xyz, box = synth.jittered_bilayer(1, G, seed=12)
box9 = np.zeros((3, 3)); box9[0, 0], box9[1, 1], box9[2, 2] = box[0]
for it in range(3):
    t0 = time.perf_counter(); r = g.tessellate_large(xyz[0], box9, 0.8, flags=3); t1 = time.perf_counter()
    print('G %d: %d atoms + %d -corr points, %d triangles, status %d, engine %d, %.1f ms with copies' % (G, xyz.shape[1], r['ncorr'], r['ntri'], r['status'], r['engine'], (t1 - t0) * 1e3), flush=True)
chk = RefOracle() if ref_available() else PortOracle()
t0 = time.perf_counter(); w = chk.tessellate(xyz, box, 0.8, 3); t1 = time.perf_counter()
print('reference: %.0f ms; rel diff area %.2e, area2d %.2e, box area equal %s' % ((t1 - t0) * 1e3, abs(r['area'] - w['area'][0]) / w['area'][0], abs(r['area2d'] - w['area2d'][0]) / w['area2d'][0], r['areabox'] == w['areabox'][0]))
