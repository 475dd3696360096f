#!/usr/bin/env python
"""bench.py -- tessellated frames/s on BASELINE.json's headline workload (cfg 3).

Workload: one leaflet of a 2,048-lipid bilayer (1,024 P atoms per frame), -corr -espace 0.8
(n = 1,156 points, 2,178 triangles per frame), 100,000 synthetic frames PER GPU per step (frames are
independent: contiguous frame ranges per rank, no collective on the data path => weak scaling).
A "step" is one pass of the hot path over the whole 100k-frame batch of every rank.

  value  frames/s with the coordinates already resident in HBM (gta_b200_tessellate_device),
         timed with CUDA events on the launching stream, max over ranks.
  e2e    the same metric through the reference-facing C-ABI with HOST buffers
         (gta_b200_tessellate_batch: pinned staging, H2D, kernel, D2H all inside the timed region).
  roofline / cpu_baseline: see DESIGN.md §5.

--impl reference times the reference's own OpenMP CPU path (oracle/_ref when it was compiled, else
the oracle port) on this box's host cores, on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NATOMS = 1024
GRID = 32
ESPACE = 0.8
FLAGS = 1                      # GTA_CORRECT
TRI_PER_FRAME = 2178           # 2(n-1) - h with n = 1156, h = 132 (reference delaunay_tri.c:608)
ALG_BYTES_PER_FRAME = NATOMS * 24 + 72 + 16      # coordinates + box in, 2 reals out
ALG_FLOP_PER_FRAME = 203.6e3                      # SURVEY.md §8d implementation-independent lower bound
FP64_PEAK_TFLOPS_NOMINAL = 37.2                   # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)
# dram read+write per frame from ncu --set full captures (profiles/): the fused CTA-per-frame kernel (6,000 frames) and the
# dominant kernel of the lane-per-frame pipeline, gt::lpf_wave_kernel, on one launch of 50,000 frames (76.0 GB read +
# 21.7 GB written: the frames' workspaces live in HBM and 50,000 of them do not fit in L2)
NCU_DRAM_BYTES_PER_FRAME_FUSED = 26254
NCU_DRAM_BYTES_PER_FRAME_LPF = 1955000
# share of the step's GPU time spent in gt::lpf_wave_kernel, from the ncu launch list of this command
# (profiles/r1_launches_summary.txt; the rest: pre-pass 8.6 %, area kernel 4.4 %)
LPF_WAVE_SHARE = 0.866
LPF_MIN_FRAMES = 32768               # launches at least this large use the lane-per-frame pipeline (gta_b200_set_batch_threshold)
WORKLOAD = "cfg3: 2,048-lipid bilayer leaflet, 1,024 atoms/frame, -corr -espace 0.8, 3-D area"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def gen_frames_device(torch, nframes: int, seed: int, device):
    """Synthetic cfg-3 frames generated on the device (same recipe as g_tessla_b200.synth.jittered_bilayer:
    32x32 jittered grid, cell 0.8 nm, jitter +-0.35 cell, box 25.6(1 + N(0, 0.005)), z = 4 + N(0, 0.15),
    all coordinates rounded to the float32 grid and kept strictly inside the box)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64, f32 = torch.float64, torch.float32
    box = torch.empty(nframes, 3, dtype=f64, device=device)
    scale = 1.0 + 0.005 * torch.randn(nframes, 2, dtype=f64, device=device, generator=g)
    box[:, :2] = (GRID * 0.8 * scale).to(f32).to(f64)
    box[:, 2] = 8.0
    idx = torch.arange(NATOMS, device=device)
    ix, iy = (idx % GRID).to(f64), (idx // GRID).to(f64)
    xyz = torch.empty(nframes, NATOMS, 3, dtype=f64, device=device)
    tiny = 1e-28
    for d, base in ((0, ix), (1, iy)):
        jit = (torch.rand(nframes, NATOMS, dtype=f64, device=device, generator=g) * 2 - 1) * 0.35
        v = ((base[None, :] + 0.5 + jit) / GRID * box[:, None, d]).to(f32).to(f64)
        hi = torch.nextafter(box[:, None, d].to(f32), torch.zeros((), dtype=f32, device=device)).to(f64)
        xyz[:, :, d] = torch.minimum(torch.clamp_min(v, tiny), hi)
        del jit, v
    xyz[:, :, 2] = (4.0 + 0.15 * torch.randn(nframes, NATOMS, dtype=f64, device=device, generator=g)).to(f32).to(f64)
    return xyz, box


def gen_frames_host(nframes: int, seed: int):
    from g_tessla_b200 import synth
    return synth.jittered_bilayer(nframes, GRID, seed=seed)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    from oracle import PortOracle, RefOracle, ref_available
    kind = "reference" if ref_available() else "port"
    orc = RefOracle() if kind == "reference" else PortOracle()
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every core this process may run on explicitly
    # (delaunay_tessellate calls omp_set_num_threads(nthreads) for nthreads > 0, reference gta_tri.c:90-92)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    xyz, box = gen_frames_host(512, seed=3)
    t = orc.tessellate(xyz, box, ESPACE, FLAGS, nthreads=cores)["seconds"]       # calibration (also warms the threads)
    per_step = max(512, min(20000, int(512 * args.ref_seconds_per_step / max(t, 1e-6))))
    xyz, box = gen_frames_host(per_step, seed=3)
    for _ in range(args.warmup):
        orc.tessellate(xyz[:max(64, per_step // 8)], box[:max(64, per_step // 8)], ESPACE, FLAGS, nthreads=cores)
    secs = [orc.tessellate(xyz, box, ESPACE, FLAGS, nthreads=cores)["seconds"] for _ in range(args.steps)]
    tot = sum(secs)
    v = per_step * args.steps / tot
    sample = "%d frames per step (bounded sample of the 100k-frame workload), %d OpenMP threads" % (per_step, cores)
    line = {"impl": "reference", "metric": "tessellated frames/sec", "value": v, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mtriangles_per_s": v * TRI_PER_FRAME / 1e6,
            "config": {"workload": WORKLOAD, "frames_per_step": per_step, "natoms": NATOMS, "points_per_frame": 1156},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def run_native(args):
    import numpy as np
    import torch
    import g_tessla_b200 as g

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = g.lib()
    F = args.frames
    pk, pk_src = peaks()

    xyz_d, box_d = gen_frames_device(torch, F, seed=3 + 1000 * rank, device=dev)
    xyz_h = torch.empty(xyz_d.shape, dtype=torch.float64, pin_memory=True).copy_(xyz_d)
    box_h = torch.empty(box_d.shape, dtype=torch.float64, pin_memory=True).copy_(box_d)
    max_points = NATOMS + 4 + 4 * int(torch.floor(box_d[:, :2].max() / ESPACE).item())
    T = g.Tessellator(F, NATOMS, max_points, ESPACE, FLAGS, device=str(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (value, roofline)
    for _ in range(args.warmup):
        T.run(xyz_d, box_d, 3)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gathered = [torch.empty_like(T.area) for _ in range(world)] if world > 1 else None
    launches_dev = 0
    for a, b in ev:
        a.record(); T.run(xyz_d, box_d, 3); b.record()
        launches_dev += L.gta_b200_last_launches()
        if world > 1:
            dist.all_gather(gathered, T.area)          # the path's only exchange: final per-frame area gather (0.8 MB per rank)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    kern_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    status_bad = int((T.status != 0).sum().item())
    area_dev = T.area.cpu().numpy().copy()

    # ---------------- end-to-end arm through the host C-ABI (pinned host buffers in, host results out)
    area = np.zeros(F); areabox = np.zeros(F); status = np.zeros(F, dtype=np.int32)

    def e2e_step():
        rc = L.gta_b200_tessellate_batch(xyz_h.data_ptr(), 3, box_h.data_ptr(), 3, F, NATOMS, ESPACE, FLAGS,
                                         local_rank, args.streams, area.ctypes.data, None, areabox.ctypes.data,
                                         status.ctypes.data, None, 0, None, None, 0, None)
        if rc != 0:
            raise SystemExit("gta_b200_tessellate_batch failed: " + L.gta_b200_last_error().decode())
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    launches_e2e = 0
    for _ in range(args.steps):
        e2e_step()
        launches_e2e += L.gta_b200_last_launches()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    if not np.allclose(area, area_dev, rtol=1e-12, atol=0) or status.any() or status_bad:
        st_dev = T.status.cpu().numpy()
        raise SystemExit("bench.py: device-resident and host-batch results differ or a frame failed: "
                         "status(dev) %s status(host) %s, %d areas differ, first %s vs %s" % (
                             np.unique(st_dev, return_counts=True), np.unique(status, return_counts=True),
                             int((area != area_dev).sum()), area[:3], area_dev[:3]))

    times = torch.tensor([total_ms, e2e_s * 1e3, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        frames_all = F * world * args.steps
        value = frames_all / (total_ms / 1e3)
        e2e_v = frames_all / (e2e_ms / 1e3)
        lpf = F >= LPF_MIN_FRAMES
        # the dominant kernel: one launch of the fused kernel per step, or `parts` launches of gt::lpf_wave_kernel (the
        # library splits a batch evenly; 3 kernels per part), which take LPF_WAVE_SHARE of the pipeline's time
        parts = max(1, launches_dev // args.steps // 3) if lpf else 1
        dom_ms = kern_ms * (LPF_WAVE_SHARE if lpf else 1.0) / parts
        frames_per_launch = F / parts
        ach = ALG_BYTES_PER_FRAME * frames_per_launch / (dom_ms / 1e3) / 1e9    # GB/s of algorithmic bytes per launch
        fl = ALG_FLOP_PER_FRAME * frames_per_launch / (dom_ms / 1e3) / 1e12
        thr, smem, fps = (g.api.C.c_int(), g.api.C.c_size_t(), g.api.C.c_int())
        L.gta_b200_launch_shape(max_points, NATOMS, g.api.C.byref(thr), g.api.C.byref(smem), g.api.C.byref(fps))
        line = {
            "metric": "tessellated frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mtriangles_per_s": value * TRI_PER_FRAME / 1e6,
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "natoms": NATOMS, "points_per_frame": 1156,
                       "l2": "inputs larger than L2 (%.2f GB of coordinates per step per GPU)" % (F * NATOMS * 24 / 1e9),
                       "path": ("lane per frame: pre-pass + phase-sorted state machine (one lane = one frame, workspace in HBM) + area kernel"
                                if lpf else "fused kernel, one CTA per frame in shared memory"),
                       "fused_kernel_shape": {"threads_per_frame": thr.value, "smem_bytes_per_frame": smem.value, "frames_per_sm": fps.value},
                       "parallelism": "frames sharded over %d GPU(s), no collective" % world},
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": F * (NATOMS * 24 + 24),
                    "d2h_bytes_per_step": F * (3 * 8 + 3 * 4), "timing": "host wall clock around the synchronous C-ABI call",
                    "streams": args.streams},
            "gpu_launches": launches_dev + launches_e2e,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "traffic": (NCU_DRAM_BYTES_PER_FRAME_LPF if lpf else NCU_DRAM_BYTES_PER_FRAME_FUSED) * frames_per_launch / 1e9,
                         "traffic_unit": "GB per launch (ncu dram bytes per frame x frames per launch)",
                         "peak_source": pk_src,
                         "kernel": ("gt::lpf_wave_kernel<512,false> (%.1f%% of the step; pre-pass gt::tessellate_kernel<256,4,true> 8.6%%, gt::lpf_area_kernel 4.4%%)" % (100 * LPF_WAVE_SHARE)) if lpf else "gt::tessellate_kernel<128,3>",
                         "launches_per_step": parts, "frames_per_launch": frames_per_launch,
                         "kernel_ms_per_launch": dom_ms, "pipeline_ms_per_step": kern_ms, "alg_bytes_per_frame": ALG_BYTES_PER_FRAME,
                         "note": "bound by the latency of dependent scattered 32-byte reads (two to three levels per iteration, L2 hit rate ~30 %); algorithmic bytes are 79x below the HBM traffic and the traffic is 12 % of HBM peak; see DESIGN.md section 5"},
            "roofline_fp64": {"achieved": fl, "peak": FP64_PEAK_TFLOPS_NOMINAL, "unit": "TFLOP/s", "frac": fl / FP64_PEAK_TFLOPS_NOMINAL,
                              "peak_source": "nominal", "alg_flop_per_frame": ALG_FLOP_PER_FRAME},
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(xyz_h.numpy(), box_h.numpy(), area_dev, args)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(xyz, box, area_gpu, args):
    """The reference's OpenMP frame loop on this box's host cores, on a bounded sample of the same
    frames the GPU just processed (also a last parity check: areas within 1e-12 relative)."""
    import numpy as np
    from oracle import PortOracle, RefOracle, ref_available
    kind = "reference" if ref_available() else "port"
    orc = RefOracle() if kind == "reference" else PortOracle()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    t = orc.tessellate(xyz[:512], box[:512], ESPACE, FLAGS, nthreads=cores)["seconds"]
    n = max(512, min(len(xyz), int(512 * args.cpu_seconds / max(t, 1e-6))))
    r = orc.tessellate(xyz[:n], box[:n], ESPACE, FLAGS, nthreads=cores)
    rel = float(np.max(np.abs(r["area"] - area_gpu[:n]) / np.abs(r["area"])))
    t1 = orc.tessellate(xyz[:256], box[:256], ESPACE, FLAGS, nthreads=1)["seconds"]
    return {"value": n / r["seconds"], "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": "first %d of the step's %d frames, one run of %.1f s, OMP threads = all (%d)" % (n, len(xyz), r["seconds"], cores),
            "one_thread_frames_per_s": 256 / t1, "max_rel_area_diff_vs_gpu": rel, "parity_ok": bool(rel <= 1e-12)}


_REAL_STDOUT = None


def own_stdout():
    """Rank 0 must print exactly ONE line on stdout. Libraries below us write banners to file descriptor 1 (NCCL's version line
    is one), so point fd 1 at stderr for the whole run and keep the real stdout for the JSON line only."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=100000, help="frames per GPU per step")
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds-per-step", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
