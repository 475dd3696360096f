"""Differential fuzz: host emulation of the device headers vs the compiled reference on nasty inputs."""
import ctypes as C, numpy as np, sys, time
sys.path.insert(0, '/root/repo')
from oracle import RefOracle
ref = RefOracle()
emu = C.CDLL('/root/repo/tests/emu/libemu_dt.so')
dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
emu.emu_triangulate.argtypes = [dp, C.c_int, ip]; emu.emu_triangulate.restype = C.c_int
emu.emu_triangulate_lpf2.argtypes = [dp, C.c_int, C.c_int, C.c_uint, ip, C.c_void_p]; emu.emu_triangulate_lpf2.restype = C.c_int
import os
LPF = os.environ.get('LPF') == '1'
def emu_tri(p):
    p = np.ascontiguousarray(p, dtype=np.float64); n = len(p)
    out = np.empty(3*2*n, dtype=np.int32)
    # LPF=1: the walker state machine, split into 2^K subtrees (K random) stepped in a random interleaving
    nt = emu.emu_triangulate_lpf2(p.ctypes.data_as(dp), n, int(rng.integers(0, 6)), int(rng.integers(0, 1 << 30)), out.ctypes.data_as(ip), None) if LPF else emu.emu_triangulate(p.ctypes.data_as(dp), n, out.ctypes.data_as(ip))
    return nt if nt < 0 else out[:3*nt].reshape(-1,3).copy()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
bad = 0; t0 = time.time()
for it in range(N):
    kind = it % 6
    if kind == 0:      # small integer grid, unique points: many collinear/cocircular
        g = int(rng.integers(2, 9)); n = int(rng.integers(2, g*g+1))
        idx = rng.choice(g*g, size=n, replace=False); p = np.stack([idx % g, idx // g], 1).astype(float)
    elif kind == 1:    # larger integer grid
        g = int(rng.integers(8, 40)); n = int(rng.integers(4, min(g*g, 600)))
        idx = rng.choice(g*g, size=n, replace=False); p = np.stack([idx % g, idx // g], 1) * 0.25
    elif kind == 2:    # rectangle boundary points + interior jitter (like -corr)
        m = int(rng.integers(1, 12)); k = int(rng.integers(0, 60))
        e = (np.arange(m) + 0.5) * 0.8; bx = m * 0.8 + rng.random() * 0.5
        p = [[0,0],[bx,0],[bx,bx],[0,bx]] + [[x,0] for x in e] + [[x,bx] for x in e] + [[0,y] for y in e] + [[bx,y] for y in e]
        p = np.array(p + (rng.random((k,2))*bx*0.98+0.01*bx).astype(np.float32).astype(float).tolist())
        p = p[rng.permutation(len(p))]
    elif kind == 3:    # hex lattice dyadic with defects
        nx, ny = int(rng.integers(2, 14)), int(rng.integers(2, 14)); q = 2.0**-6
        j, i = np.divmod(np.arange(nx*ny), nx)
        p = np.stack([(i + 0.5*(j%2))*32*q, j*28*q], 1)
        m = rng.random(len(p)) < 0.15; p[m] += rng.integers(-3, 4, size=(int(m.sum()), 2))*q
        p = np.unique(p, axis=0); p = p[rng.permutation(len(p))]
    elif kind == 4:    # points on few lines / circles
        n = int(rng.integers(3, 80)); t = rng.integers(-20, 21, n)
        p = np.stack([t, (t*int(rng.integers(-3,4))) + rng.integers(0, 2, n)*int(rng.integers(1, 5))], 1).astype(float)
        p = np.unique(p, axis=0); p = p[rng.permutation(len(p))]
    else:              # random general position
        n = int(rng.integers(2, 300)); p = rng.random((n, 2))
    if len(p) < 2: continue
    a, nv = ref.triangulate(p)
    b = emu_tri(p)
    ok = isinstance(b, np.ndarray) and a.shape == b.shape and (a == b).all()
    if not ok:
        bad += 1
        if bad <= 5:
            print('MISMATCH kind', kind, 'n', len(p), 'emu', b if not isinstance(b, np.ndarray) else b.shape, 'ref', a.shape)
            np.save('/tmp/fuzz_bad_%d.npy' % bad, p)
print('cases', N, 'bad', bad, '%.1fs' % (time.time() - t0))
