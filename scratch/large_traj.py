"""A trajectory of large frames (G x G lipids each, -corr -2d) through the kept delaunay_tessellate: frames/s on 1..N GPUs
(GTA_B200_NGPUS), results compared between the runs and with the reference on the first frames.  python scratch/large_traj.py [G] [frames]"""
import ctypes as C, os, subprocess, sys, time, numpy as np
sys.path.insert(0, '.')
G = int(sys.argv[1]) if len(sys.argv) > 1 else 316
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
code = r'''
import ctypes as C, sys, time, numpy as np, faulthandler
faulthandler.enable()
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
G, F = int(sys.argv[1]), int(sys.argv[2])
xyz, box = synth.jittered_bilayer(F, G, seed=5)
L = g.lib(); dp = C.POINTER(C.c_double)
class TriArea(C.Structure):
    _fields_ = [("area", dp), ("area2D", dp), ("area2Dbox", dp), ("natoms", C.c_int), ("nframes", C.c_int)]
frames = [np.ascontiguousarray(xyz[f]) for f in range(F)]
xptr = (dp * F)(*[fr.ctypes.data_as(dp) for fr in frames])
box9 = np.zeros((F, 3, 3)); box9[:, 0, 0] = box[:, 0]; box9[:, 1, 1] = box[:, 1]; box9[:, 2, 2] = box[:, 2]
L.delaunay_tessellate.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.POINTER(TriArea), C.c_ubyte]
for it in range(4):
    a = TriArea(None, None, None, xyz.shape[1], F)
    t0 = time.perf_counter(); L.delaunay_tessellate(xptr, box9.ctypes.data, 0.8, 0, C.byref(a), 3); t1 = time.perf_counter()
    area = np.ctypeslib.as_array(a.area, shape=(F,)).copy(); L.free_tri_area(C.byref(a))
print("RESULT %.3f %s" % (F / (t1 - t0), " ".join(repr(float(v)) for v in area)))
'''
res = {}
import torch
for ng in [n for n in (1, 2, 4, 8) if n <= torch.cuda.device_count()]:
    out = subprocess.run([sys.executable, "-c", code, str(G), str(F)], env=dict(os.environ, GTA_B200_NGPUS=str(ng)), capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    if not line: print("rc", out.returncode, out.stdout[-1500:], out.stderr[-3000:]); sys.exit(1)
    tok = line[0].split(); res[ng] = (float(tok[1]), np.array([float(v) for v in tok[2:]]))
    print("%d GPU(s): %d frames of %d atoms, %.1f frames/s" % (ng, F, G * G, res[ng][0]), flush=True)
base = res[min(res)][1]
for ng, (_, a) in res.items(): assert np.array_equal(a, base), ng
from g_tessla_b200 import synth
from oracle import RefOracle, PortOracle, ref_available
chk = RefOracle() if ref_available() else PortOracle()
xyz, box = synth.jittered_bilayer(F, G, seed=5)
t0 = time.perf_counter(); w = chk.tessellate(xyz[:2], box[:2], 0.8, 3); t1 = time.perf_counter()
print("reference: %.2f frames/s on one thread; max rel diff %.2e; identical across GPU counts" % (2 / (t1 - t0), float(np.max(np.abs(base[:2] - w["area"]) / w["area"]))))
