"""CPU, world_size 2 over gloo: the frame-sharding / final-gather logic of the multi-GPU path.
The per-rank compute is a stand-in (the oracle port) because this machine has no GPU; the GPU runs use the same
partition with the CUDA library (bench.py under torchrun, gta_b200_tessellate_multi in-process)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_range_partition():
    from g_tessla_b200.dist import frame_range
    for F in (0, 1, 7, 100, 100000):
        for W in (1, 2, 3, 8):
            r = [frame_range(F, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == F
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))          # contiguous, disjoint, complete
            assert max(b - a for a, b in r) <= (F + W - 1) // W


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from g_tessla_b200 import synth
    from g_tessla_b200.dist import tessellate_sharded
    from oracle import PortOracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port_oracle = PortOracle()

    def compute(x, b, e, f):
        r = port_oracle.tessellate(x, b, e, f)
        r["status"] = np.zeros(len(x), dtype=np.int32)
        return r
    c = synth.config(2, nframes=37)                         # odd count: ranks get 19 and 18 frames
    out = tessellate_sharded(c["xyz"], c["box"], c["espace"], c["flags"], compute=compute)
    dist.destroy_process_group()
    q.put((rank, out["area"], out["area2d"], out["areabox"]))


def test_world2_sharded_equals_unsharded():
    import torch.multiprocessing as mp
    from g_tessla_b200 import synth
    from oracle import PortOracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=180) for _ in range(2)]
    [p.join(timeout=60) for p in ps]
    c = synth.config(2, nframes=37)
    want = PortOracle().tessellate(c["xyz"], c["box"], c["espace"], c["flags"])
    for rank, area, area2d, areabox in res:
        assert np.array_equal(area, want["area"]) and np.array_equal(area2d, want["area2d"]) and np.array_equal(areabox, want["areabox"])
