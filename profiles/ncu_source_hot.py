"""Aggregate an ncu `--page source --csv --print-source sass,cuda` dump by source file:line.
usage: ncu_source_hot.py dump.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.OrderedDict(); cur_file = None; hdr = None; func = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1][:60]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) // 2: continue
    d = dict(zip(hdr, r))
    # rows of the CUDA-source view have a Line No; SASS rows have an Address
    ln = d.get("Line No", "")
    if not ln.isdigit(): continue
    try:
        ie = float(d.get("Instructions Executed", "0") or 0); te = float(d.get("Thread Instructions Executed", "0") or 0)
        smp = float(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    k = (cur_file, int(ln)); a = agg.setdefault(k, [0, 0, 0, d.get("Source", "")[:90]])
    a[0] += ie; a[1] += te; a[2] += smp
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print("total warp-inst %.3g, samples %d" % (tot_i, tot_s))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%5.1f%% smp %5.1f%% inst  lanes %4.1f  %s:%d  %s" % (100 * a[2] / max(tot_s, 1), 100 * a[0] / max(tot_i, 1), a[1] / max(a[0], 1), f, ln, a[3].strip()))
