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
  strong_100k  BASELINE configs[2] literally: ONE 100,000-frame trajectory sharded over the GPUs (strong scaling).
  kept_api     the drop-in entry point's engine (gta_b200_tessellate_frames) on per-frame heap blocks, all GPUs, from rank 0.
  roofline / cpu_baseline: see DESIGN.md §5 (kernel times and the FP64 peak are measured by the run that prints them).

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
FP64_PEAK_TFLOPS_NOMINAL = 37.2                   # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (fallback if the DFMA micro-benchmark fails)
STRONG_FRAMES = 100000                            # BASELINE.json configs[2]: ONE 100k-frame trajectory, sharded over the GPUs
TRAFFIC_JSON = os.path.join("profiles", "r2_star_traffic.json")   # written by profiles/ncu_keys.py --json from an ncu --set full capture
WORKLOAD = "cfg3: 2,048-lipid bilayer leaflet, 1,024 atoms/frame, -corr -espace 0.8, 3-D area"


def workload_config(frames_per_gpu):
    """Identical in both arms (the native one and --impl reference), so that the driver's comparison sees one config."""
    return {"workload": WORKLOAD, "frames_per_gpu_per_step": frames_per_gpu, "natoms": NATOMS, "points_per_frame": 1156,
            "espace": ESPACE, "flags": "GTA_CORRECT", "headline": "weak scaling: every GPU tessellates its own frames_per_gpu_per_step frames; "
            "the 100k-frame trajectory of BASELINE configs[2] sharded over the GPUs (strong scaling) is reported beside it as strong_100k"}


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception:
        return ""


def kernel_sources_sha():
    """Hash of the CUDA sources the captured kernel is built from: the traffic figure is stale when they change."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "g_tessla_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


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
            "config": workload_config(args.frames),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def run_native(args):
    import ctypes as C
    import numpy as np
    import torch
    import g_tessla_b200 as g
    from g_tessla_b200.dist import frame_range, gather_frames

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")          # host-side waits that must not put a spinning kernel on the GPUs
    L = g.lib()
    F = args.frames
    pk, pk_src = peaks()

    xyz_d, box_d = gen_frames_device(torch, F, seed=3 + 1000 * rank, device=dev)
    xyz_h = torch.empty(xyz_d.shape, dtype=torch.float64, pin_memory=True).copy_(xyz_d)
    box_h = torch.empty(box_d.shape, dtype=torch.float64, pin_memory=True).copy_(box_d)
    max_points = NATOMS + 4 + 4 * int(torch.floor(box_d[:, :2].max() / ESPACE).item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_arm(nf, gather_total=None):
        """nf of this rank's frames, already in HBM: (ms per step over the timed region, ms inside the library call, launches, Tessellator)"""
        T = g.Tessellator(nf, NATOMS, max_points, ESPACE, FLAGS, device=str(dev))
        x, bx = xyz_d[:nf], box_d[:nf]
        for _ in range(args.warmup):
            T.run(x, bx, 3)
        barrier()
        g.api.stage_ms(local_rank, reset=True)
        g.api.star_counters(local_rank, reset=True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        e0.record()
        for a, b in ev:
            a.record(); T.run(x, bx, 3); b.record()
            launches += L.gta_b200_last_launches()
            if world > 1:                                     # the path's only exchange: the final per-frame area gather
                gather_frames(T.area, gather_total if gather_total else nf * world)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), sum(a.elapsed_time(b) for a, b in ev) / args.steps, launches, T

    def e2e_arm(nf):
        """the same frames through the reference-facing C-ABI with HOST buffers: H2D, kernels, D2H inside the timed region"""
        area = np.zeros(nf); areabox = np.zeros(nf); status = np.zeros(nf, dtype=np.int32)

        def step():
            rc = L.gta_b200_tessellate_batch(xyz_h.data_ptr(), 3, box_h.data_ptr(), 3, nf, NATOMS, ESPACE, FLAGS,
                                             local_rank, args.streams, area.ctypes.data, None, areabox.ctypes.data,
                                             status.ctypes.data, None, 0, None, None, 0, None)
            if rc != 0:
                raise SystemExit("gta_b200_tessellate_batch failed: " + L.gta_b200_last_error().decode())
        for _ in range(args.warmup):
            step()
        barrier()
        t0 = time.perf_counter()
        launches = 0
        for _ in range(args.steps):
            step()
            launches += L.gta_b200_last_launches()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, launches, area, status

    # ---------------- weak arm (the headline): every rank its own F frames
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    total_ms, kern_ms, launches_dev, T = resident_arm(F)
    stages = g.api.stage_ms(local_rank)
    certified, fallback = g.api.star_counters(local_rank)
    status_bad = int((T.status != 0).sum().item())
    area_dev = T.area.cpu().numpy().copy()
    e2e_ms, launches_e2e, area, status = e2e_arm(F)
    clk = clocks.stop() if rank == 0 else None
    if not np.allclose(area, area_dev, rtol=1e-12, atol=0) or status.any() or status_bad:
        st_dev = T.status.cpu().numpy()
        raise SystemExit("bench.py: device-resident and host-batch results differ or a frame failed: "
                         "status(dev) %s status(host) %s, %d areas differ, first %s vs %s" % (
                             np.unique(st_dev, return_counts=True), np.unique(status, return_counts=True),
                             int((area != area_dev).sum()), area[:3], area_dev[:3]))
    del T

    # ---------------- strong arm: ONE trajectory of STRONG_FRAMES frames, contiguous ranges per rank (BASELINE configs[2])
    lo, hi = frame_range(args.strong_frames, rank, world)
    nf_s = hi - lo
    if world == 1 and nf_s == F:
        s_total_ms, s_e2e_ms, s_launches = total_ms, e2e_ms, 0          # the same measurement
    else:
        s_total_ms, _, l1, Ts = resident_arm(nf_s, gather_total=args.strong_frames)
        s_e2e_ms, l2, s_area, s_status = e2e_arm(nf_s)
        if s_status.any() or not np.allclose(s_area, area_dev[:nf_s], rtol=1e-12, atol=0):
            raise SystemExit("bench.py: strong-scaling arm disagrees with the weak arm on the same frames")
        s_launches = l1 + l2
        del Ts

    times = torch.tensor([total_ms, e2e_ms, kern_ms, s_total_ms, s_e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms, s_total_ms, s_e2e_ms = [float(x) for x in times.tolist()]

    # ---------------- the drop-in entry point (delaunay_tessellate -> gta_b200_tessellate_frames) on every GPU of the box:
    # rank 0 alone, per-frame heap blocks as read_traj leaves them (extern/gkut/src/gkut_io.c:37); the other ranks wait on the host
    kept = None
    if rank == 0:
        kept = kept_api_arm(L, C, np, xyz_h, box_h, area_dev, world, args)
    if world > 1:
        dist.barrier(group=host_group)

    if rank == 0:
        frames_all = F * world * args.steps
        value = frames_all / (total_ms / 1e3)
        e2e_v = frames_all / (e2e_ms / 1e3)
        nl = max(1, stages["launches"])
        # dominant kernel of the step, per launch, measured in this run: the star kernel (or the walker kernel when the stars were skipped)
        dom = "star" if stages["star"] >= stages["walk"] else "walk"
        dom_ms = stages[dom] / nl
        frames_per_launch = F * args.steps / nl
        ach = ALG_BYTES_PER_FRAME * frames_per_launch / (dom_ms / 1e3) / 1e9     # GB/s of algorithmic bytes per launch
        stage_sum = max(1e-9, stages["star"] + stages["prepass"] + stages["walk"] + stages["area"])
        # ncu dram bytes per frame of that kernel: from the committed capture summary, flagged when the kernels changed since
        traffic, traffic_meta, ncu_extra = None, {"source": TRAFFIC_JSON, "stale": None}, None
        try:
            tj = json.load(open(os.path.join(ROOT, TRAFFIC_JSON)))
            traffic = tj["dram_bytes_per_frame"] * frames_per_launch / 1e9
            ncu_extra = {k: tj.get(k) for k in ("fp64_pipe_active_pct", "issue_active_pct", "active_lanes_per_instruction", "warp_instructions_per_frame", "registers_per_thread")}
            traffic_meta.update({"dram_bytes_per_frame": tj["dram_bytes_per_frame"], "captured_frames": tj.get("frames"),
                                 "capture_git": tj.get("git"), "stale": tj.get("sources_sha") != kernel_sources_sha()})
        except Exception as e:
            traffic_meta["error"] = str(e)
        fp64_peak = L.gta_b200_fp64_peak_tflops(local_rank)
        fp64_src = "measured (DFMA micro-benchmark, gta_b200_fp64_peak_tflops)"
        if not fp64_peak > 0:
            fp64_peak, fp64_src = FP64_PEAK_TFLOPS_NOMINAL, "nominal"
        fl = ALG_FLOP_PER_FRAME * value / world / 1e12
        line = {
            "metric": "tessellated frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mtriangles_per_s": value * TRI_PER_FRAME / 1e6,
            "config": workload_config(F),
            "run": {"l2": "inputs larger than L2 (%.2f GB of coordinates per step per GPU)" % (F * NATOMS * 24 / 1e9),
                    "path": "star engine: one fused launch per chunk (load, -corr, grid binning, per-point Delaunay stars with exact predicates, areas); "
                            "frames it does not certify (cocircular ties) go through pre-pass + walker kernel + area kernel inside the same call",
                    "frames_certified_by_stars": certified, "frames_through_divide_and_conquer": fallback,
                    "parallelism": "frames sharded over %d GPU(s), no collective on the data path (final per-frame area gather only)" % world,
                    "generator": "device (torch CUDA generator), recipe of g_tessla_b200.synth.jittered_bilayer; seed 3 + 1000 * rank"},
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": F * (NATOMS * 24 + 24),
                    "d2h_bytes_per_step": F * (3 * 8 + 3 * 4), "timing": "host wall clock around the synchronous C-ABI call",
                    "streams": args.streams},
            "strong_100k": {"frames_total": args.strong_frames, "frames_per_gpu": nf_s, "scaling": "strong",
                            "value": args.strong_frames * args.steps / (s_total_ms / 1e3), "e2e": args.strong_frames * args.steps / (s_e2e_ms / 1e3),
                            "unit": "frames/s", "ms_per_step": s_total_ms / args.steps,
                            "note": "one 100k-frame trajectory, contiguous frame ranges per GPU (g_tessla_b200.dist.frame_range), final area gather inside the timed region"},
            "kept_api": kept,
            "gpu_launches": launches_dev + launches_e2e + s_launches + (kept or {}).get("launches", 0),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read + dram__bytes_write per frame x frames per launch)",
                         "traffic_meta": traffic_meta, "peak_source": pk_src,
                         "kernel": "gt::star_kernel<256,2,false>" if dom == "star" else "gt::lpf_wave_kernel<512,false>",
                         "kernel_share_of_step": stages[dom] / stage_sum,
                         "stage_ms_per_launch": {"star": stages["star"] / nl, "prepass": stages["prepass"] / nl, "walk": stages["walk"] / nl, "area": stages["area"] / nl},
                         "launches_per_step": nl / args.steps, "frames_per_launch": frames_per_launch,
                         "kernel_ms_per_launch": dom_ms, "pipeline_ms_per_step": kern_ms, "alg_bytes_per_frame": ALG_BYTES_PER_FRAME,
                         "timing": "CUDA events on the launching stream around each kernel, inside the library (gta_b200_stage_ms4), over the timed steps",
                         "note": "the star kernel reads each coordinate once (DRAM traffic ~ the algorithmic bytes) and is bound by instruction issue and the FP64 pipe, "
                                 "not by HBM: ~2,300 double-precision predicate operations per point against 24 bytes of input (DESIGN.md section 5)"},
            "roofline_fp64": {"achieved": fl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / fp64_peak,
                              "peak_source": fp64_src, "alg_flop_per_frame": ALG_FLOP_PER_FRAME,
                              "ncu_capture_of_the_kernel": ncu_extra},
            "clocks": clk,
        }
        if ncu_extra and ncu_extra.get("warp_instructions_per_frame") and clk and clk.get("sm_mhz"):
            # what does bound the kernel: warp instructions issued, against one per cycle per SM sub-partition (4 per SM).
            # Instructions per frame from the committed ncu capture of the same sources, rate and clock from this run.
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            issue_peak = sms * 4 * clk["sm_mhz"] * 1e6
            issue_ach = ncu_extra["warp_instructions_per_frame"] * frames_per_launch / (dom_ms / 1e3)
            line["roofline_issue"] = {"achieved": issue_ach / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s", "frac": issue_ach / issue_peak,
                                      "active_lanes_per_instruction": ncu_extra.get("active_lanes_per_instruction"),
                                      "note": "the kernel's own bound (DESIGN.md section 5): issue slots x lanes; instructions per frame from the ncu capture, stale with it"}
        if world == 1 and not args.no_large_frames:
            line["large_frames"] = large_frames_arm(g, np)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(xyz_h.numpy(), box_h.numpy(), area_dev, args)
            if kept and line["cpu_baseline"].get("value"):
                kept["vs_cpu"] = kept["frames_per_s"] / line["cpu_baseline"]["value"]
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def large_frames_arm(g, np):
    """Outside the headline metric, for the record (BASELINE configs[3] and the large-frame form of the kept entry points):
    ONE frame of 10^6 points through gta_b200_triangulate_large (star engine, triangle list to the host) and one frame of
    10^6 atoms with -corr through gta_b200_tessellate_large. Host buffers in, host results out; the second call of each is
    timed. Never fails the bench: an error is reported in the object."""
    from g_tessla_b200 import synth
    out = {}
    try:
        p = synth.uniform_patch(1_000_000, seed=4)
        old = g.set_path(g.PATH_STAR)
        try:
            g.triangulate_large(p)
            t0 = time.perf_counter(); tri, st, ms = g.triangulate_large(p); t1 = time.perf_counter()
        finally:
            g.set_path(old)
        out["cfg4_one_frame_1e6_points"] = {"triangles": int(len(tri)), "status": int(st), "engine": int(g.lib().gta_b200_large_last_engine()),
                                            "device_ms": float(ms), "host_to_host_ms": (t1 - t0) * 1e3,
                                            "note": "star engine: the reference's triangle set (tests/test_gpu_large.py::test_cfg4_star_engine), in cell order"}
        xyz, box = synth.jittered_bilayer(1, 1000, seed=12)
        b9 = np.zeros((3, 3)); b9[0, 0], b9[1, 1], b9[2, 2] = box[0]
        g.tessellate_large(xyz[0], b9, 0.8, flags=3)
        t0 = time.perf_counter(); r = g.tessellate_large(xyz[0], b9, 0.8, flags=3); t1 = time.perf_counter()
        out["one_frame_1e6_atoms_corr_2d"] = {"ncorr": int(r["ncorr"]), "triangles": int(r["ntri"]), "status": int(r["status"]), "engine": int(r["engine"]),
                                              "host_to_host_ms": (t1 - t0) * 1e3}
    except Exception as e:                                              # noqa: BLE001 -- reported, not raised
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out


def kept_api_arm(L, C, np, xyz_h, box_h, area_dev, ngpus, args):
    """delaunay_tessellate's own data layout through the kept C entry point's engine: x[fr] = one heap block per frame,
    box[fr] = 3x3 matrix, results into host arrays; ngpus devices driven from this one process (one host thread each)."""
    F = xyz_h.shape[0]
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
    nbytes = NATOMS * 24
    base = xyz_h.data_ptr()
    ptrs = (C.c_void_p * F)()
    for f in range(F):
        p = libc.malloc(nbytes)
        C.memmove(p, base + f * nbytes, nbytes)
        ptrs[f] = p
    box9 = np.zeros((F, 3, 3)); bh = box_h.numpy()
    box9[:, 0, 0] = bh[:, 0]; box9[:, 1, 1] = bh[:, 1]; box9[:, 2, 2] = bh[:, 2]
    area = np.zeros(F); areabox = np.zeros(F); status = np.zeros(F, dtype=np.int32)

    def step():
        rc = L.gta_b200_tessellate_frames(ptrs, box9.ctypes.data, F, NATOMS, ESPACE, FLAGS, ngpus, 0,
                                          area.ctypes.data, None, areabox.ctypes.data, status.ctypes.data)
        if rc != 0:
            raise SystemExit("gta_b200_tessellate_frames failed: " + L.gta_b200_last_error().decode())
    step()                                                     # warm-up: pinned staging buffers, workspaces on every device
    reps = max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    launches = 0
    for _ in range(reps):
        step()
        launches += L.gta_b200_last_launches()
    dt = time.perf_counter() - t0
    for f in range(F):
        libc.free(ptrs[f])
    rel = float(np.max(np.abs(area - area_dev) / np.abs(area_dev)))
    return {"entry": "gta_b200_tessellate_frames (engine of delaunay_tessellate), per-frame heap blocks", "frames": F, "ngpus": ngpus,
            "frames_per_s": F * reps / dt, "ms_per_call": 1e3 * dt / reps, "launches": launches,
            "max_rel_vs_device_arm": rel, "status_ok": bool(not status.any())}


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
    ap.add_argument("--strong-frames", type=int, default=STRONG_FRAMES, help="frames of the ONE trajectory the strong-scaling arm shards over the GPUs")
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds-per-step", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-frames", action="store_true", help="skip the two large-frame measurements reported beside the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
