"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log:
kernels of the LAST bench step only (from the last umi_pack_kernel launch of the device-resident arm back to ... see
--from-last), grouped by kernel name.  usage: launch_list.py raw.csv [marker-kernel-substring]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names = rows[hdr]
ki, mi, vi, ui, idi = names.index("Kernel Name"), names.index("Metric Name"), names.index("Metric Value"), names.index("Metric Unit"), names.index("ID")
launches = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(names):
        continue
    d = launches.setdefault(int(r[idi]), {"name": r[ki]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        d["us"] = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    else:
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        d["rd" if "read" in r[mi] else "wr"] = v * scale
ls = list(launches.values())
marker = sys.argv[2] if len(sys.argv) > 2 else "umi_pack_kernel"
# the last two steps are the e2e arm (host push) and ... keep it simple: take the kernels from the LAST marker launch on
last = max(i for i, l in enumerate(ls) if marker in l["name"])
# bench order: device-resident steps first, then the e2e arm; pick the step that starts at the FIRST marker after warm-up
firsts = [i for i, l in enumerate(ls) if marker in l["name"]]
start = firsts[1] if len(firsts) > 1 else firsts[0]
end = firsts[2] if len(firsts) > 2 else len(ls)
step = ls[start:end]
tot = sum(l["us"] for l in step)
agg = OrderedDict()
for l in step:
    nm = l["name"].split("(")[0]
    a = agg.setdefault(nm, [0.0, 0, 0.0, 0.0])
    a[0] += l["us"]; a[1] += 1; a[2] += l.get("rd", 0.0); a[3] += l.get("wr", 0.0)
print(f"# launches {start}..{end - 1} of {len(ls)}: one device-resident step; total kernel time {tot / 1000:.2f} ms, {len(step)} launches")
print("#        us  share    n  DRAM rd MB  DRAM wr MB     GB/s  kernel")
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{a[0]:11.1f} {100 * a[0] / tot:5.1f}% {a[1]:4d} {a[2]:11.1f} {a[3]:11.1f} {(a[2] + a[3]) / a[0] * 1e3 if a[0] else 0:8.1f}  {nm[:90]}")
