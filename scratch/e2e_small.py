"""End-to-end (host buffers) throughput of the batch and the per-frame-pointer entry points at small batch sizes."""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import g_tessla_b200 as g
from g_tessla_b200 import synth
L = g.lib()
sizes = [int(x) for x in sys.argv[1].split(',')]
base = synth.config(3, nframes=2000)
for nf in sizes:
    reps = (nf + 1999) // 2000
    xyz = torch.empty((nf, 1024, 3), dtype=torch.float64, pin_memory=True).copy_(torch.from_numpy(np.tile(base['xyz'], (reps, 1, 1))[:nf]))
    box = torch.empty((nf, 3), dtype=torch.float64, pin_memory=True).copy_(torch.from_numpy(np.tile(base['box'], (reps, 1))[:nf]))
    area = np.zeros(nf); ab = np.zeros(nf); st = np.zeros(nf, dtype=np.int32)
    def batch():
        rc = L.gta_b200_tessellate_batch(xyz.data_ptr(), 3, box.data_ptr(), 3, nf, 1024, 0.8, 1, 0, 4, area.ctypes.data, None, ab.ctypes.data, st.ctypes.data, None, 0, None, None, 0, None)
        assert rc == 0
    frames = [np.array(xyz[f].numpy()) for f in range(nf)]
    ptrs = (C.c_void_p * nf)(*[f.ctypes.data for f in frames])
    box9 = np.zeros((nf, 3, 3)); b = box.numpy(); box9[:, 0, 0] = b[:, 0]; box9[:, 1, 1] = b[:, 1]
    def fr():
        rc = L.gta_b200_tessellate_frames(ptrs, box9.ctypes.data, nf, 1024, 0.8, 1, 1, 0, area.ctypes.data, None, ab.ctypes.data, st.ctypes.data)
        assert rc == 0
    for name, fn in (('batch', batch), ('frames', fr)):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(4): fn()
        dt = (time.perf_counter() - t0) / 4
        print('%s chunk_min=%s frames %d: %.1f ms -> %.0f frames/s, bad %d' % (name, os.environ.get('GTA_B200_CHUNK_MIN'), nf, dt * 1e3, nf / dt, int((st != 0).sum())), flush=True)
