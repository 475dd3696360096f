"""Print the handful of ncu raw-page metrics the DESIGN/roofline notes quote (per captured launch).

  python profiles/ncu_keys.py REPORT.ncu-rep                         print the metrics
  python profiles/ncu_keys.py REPORT.ncu-rep --json OUT.json FRAMES  also write the DRAM traffic per frame of the first captured
      launch (FRAMES = frames that launch processed) with the hash of the CUDA sources and the git commit it was taken on:
      bench.py reads it for roofline.traffic and flags it stale when the sources have changed since.
"""
import csv, hashlib, json, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            print("%-70s %-12s %s" % (k, units[hdr.index(k)], d[k]))
    top = sorted(((float(d[h].replace(",", "")), h) for h in stall if d[h] not in ("", "n/a")), reverse=True)[:6]
    for v, h in top:
        print("  stall %-45s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    print("-" * 40)

if "--json" in sys.argv:
    k = sys.argv.index("--json")
    out, frames = sys.argv[k + 1], int(sys.argv[k + 2])
    d = dict(zip(hdr, rows[2]))
    def gb(name):
        v = float(d[name].replace(",", "")); u = units[hdr.index(name)].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[u]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha1()
    src = os.path.join(root, "g_tessla_b200", "csrc")
    for f in sorted(os.listdir(src)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(src, f), "rb").read())
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    git = subprocess.run(["git", "-C", root, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    def num(name):
        try:
            return float(d[name].replace(",", ""))
        except Exception:
            return None
    json.dump({"report": os.path.basename(rep), "kernel": d.get("Kernel Name", ""), "frames": frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_frame": (rd + wr) / frames, "gpu_time_ms": float(d["gpu__time_duration.sum"].replace(",", "")),
               "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "active_lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "warp_instructions_per_frame": (num("smsp__inst_executed.sum") or 0.0) / frames,
               "registers_per_thread": num("launch__registers_per_thread"),
               "sources_sha": h.hexdigest(), "git": git}, open(out, "w"), indent=1)
    print("wrote", out)
