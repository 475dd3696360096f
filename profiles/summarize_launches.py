"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = OrderedDict()
for r in rows:
    name = re.sub(r"<.*", "", r[ki].replace("void ", ""))[:60]
    m = re.search(r"gt::(star_kernel|ls_\w+|tessellate_kernel|prepass_kernel|lpf_wave_kernel|lpf_area_kernel|dfma_peak_kernel|large_\w+)", r[ki])
    if m:
        t = re.search(r"<[^>]*>", r[ki])
        name = "gt::" + m.group(1) + (t.group(0) if t else "") + " grid" + r[gi] + " block" + r[bi]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("%-90s %6s %14s %7s" % ("kernel", "count", "total_ns", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-90s %6d %14.0f %6.2f%%" % (k, a[0], a[1], 100 * a[1] / tot))
