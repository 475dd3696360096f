"""End-to-end throughput of the kept entry point's layout: one heap block per frame (gta_b200_tessellate_frames)."""
import ctypes as C, sys, time, numpy as np
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
L = g.lib(); dp = C.POINTER(C.c_double)
nf = int(sys.argv[1]); base = synth.config(3, nframes=2000)
frames = [np.ascontiguousarray(base['xyz'][f % 2000]).copy() for f in range(nf)]
box9 = np.zeros((nf, 9)); box9[:, 0] = np.tile(base['box'][:, 0], nf // 2000 + 1)[:nf]; box9[:, 4] = np.tile(base['box'][:, 1], nf // 2000 + 1)[:nf]
xp = (dp * nf)(*[fr.ctypes.data_as(dp) for fr in frames])
area = np.zeros(nf); abox = np.zeros(nf); st = np.zeros(nf, dtype=np.int32)
L.gta_b200_tessellate_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
for nthreads in (1, 0):
    for it in range(2):
        t0 = time.perf_counter()
        rc = L.gta_b200_tessellate_frames(xp, box9.ctypes.data, nf, 1024, 0.8, 1, 1, nthreads, area.ctypes.data, None, abox.ctypes.data, st.ctypes.data)
        t1 = time.perf_counter()
    print('frames API nthreads=%d: %d frames in %.1f ms -> %.0f frames/s rc=%d bad=%d' % (nthreads, nf, (t1 - t0) * 1e3, nf / (t1 - t0), rc, int((st != 0).sum())))
want = g.tessellate(base['xyz'][:100], base['box'][:100], 0.8, 1)
print('max rel', float(np.max(np.abs(area[:100] - want['area']) / want['area'])))
