"""Regenerates profiles/ncu_traffic.json (what bench.py copies into roofline*.traffic) from an ncu launch list of the committed
code:   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv \\
            python bench.py --workload C5 --steps 1 --warmup 1 --no-extras --no-cpu-baseline
usage: make_traffic.py WORKLOAD X.csv COMMIT [WORKLOAD2 Y.csv ...]   -> merges into profiles/ncu_traffic.json
DRAM bytes of ONE device-resident step, summed per stage (kernel name -> stage)."""
import csv
import json
import os
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "ncu_traffic.json")
STAGE = [("seg_", "sort"), ("radix_", "neighbours"), ("hamming_", "neighbours"), ("onehot", "neighbours"), ("expand_blocks", "neighbours"),
         ("build_items", "neighbours"), ("tile_summary", "neighbours"), ("small_buckets", "neighbours"), ("mi_", "neighbours"),
         ("umi_pack", "pack"), ("build_keys", "keys"), ("unique_", "unique"), ("HeadFlag", "unique"), ("label_sweep", "cluster"), ("uf_", "cluster"),
         ("contract_edges", "cluster"), ("expand_labels", "cluster"), ("keep_from_label", "cluster"), ("frontier", "cluster"), ("csr_", "cluster"),
         ("mark_kept", "emit"), ("Bitmap", "emit")]


def step_launches(path, marker="umi_pack_kernel"):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    ki, mi, vi, ui, idi = names.index("Kernel Name"), names.index("Metric Name"), names.index("Metric Value"), names.index("Metric Unit"), names.index("ID")
    launches = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) != len(names):
            continue
        d = launches.setdefault(int(r[idi]), {"name": r[ki], "bytes": 0.0, "us": 0.0})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            d["us"] = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1000.0)
        else:
            d["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    ls = list(launches.values())
    firsts = [i for i, l in enumerate(ls) if marker in l["name"]]
    start = firsts[1] if len(firsts) > 1 else firsts[0]          # the first step after the warm-up step
    end = firsts[2] if len(firsts) > 2 else len(ls)
    return ls[start:end]


def main():
    args = sys.argv[1:]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    out = {k: v for k, v in out.items() if k.startswith("C") or k == "_comment"}
    out["_comment"] = ("DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of ONE device-resident step per stage, from an ncu launch list "
                       "of the commit named in each entry (profiles/tools/make_traffic.py).  bench.py copies these into roofline*.traffic.")
    while args:
        wl, path, commit = args[0], args[1], args[2]
        args = args[3:]
        agg = {}
        for l in step_launches(path):
            st = next((s for pat, s in STAGE if pat in l["name"]), "other")
            a = agg.setdefault(st, {"bytes": 0.0, "us": 0.0, "launches": 0})
            a["bytes"] += l["bytes"]; a["us"] += l["us"]; a["launches"] += 1
        out[wl] = {st: {"bytes": int(a["bytes"]), "kernel_us_under_ncu": round(a["us"], 1), "launches": a["launches"],
                        "source": f"{os.path.basename(path)} @ {commit}"} for st, a in agg.items()}
    json.dump(out, open(OUT, "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
