"""A/B of run-time variants selected by environment variables, on one GPU, inputs resident in HBM.
usage: ab_env.py --workload C2 [--scale 1.0] --steps 5 VAR=a VAR=b ...   (each VAR=value is one arm; 'base' = nothing set)
Prints one JSON line per arm with the median stage times (ms) and the kept count (must agree between arms)."""
import argparse
import json
import os
import statistics
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "umi-collapse-rs_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("arms", nargs="+")
a = ap.parse_args()

import torch          # noqa: E402
import umigpu         # noqa: E402
from umigpu import synth   # noqa: E402

d, cfg = synth.generate_config(a.workload, device="cuda", scale=a.scale)
torch.cuda.synchronize()
algo = {"dir": 0, "adj": 1, "adj-upstream": 2, "cc": 3}[cfg["algo"]]
for arm in a.arms:
    sets = [] if arm == "base" else [kv.split("=", 1) for kv in arm.split(",")]
    for k, v in sets:
        os.environ[k] = v
    with umigpu.Context(cfg["umi_len"], cfg["k"], 0.5, algo, umigpu.MERGE_AVGQUAL, 0) as ctx:
        acc = {}
        for it in range(a.steps + 2):
            ctx.reset()
            ctx.push_reads(d["tid"], d["pos"], d["rev"], d["umi"], d["score"], sync=False)
            ctx.run()
            if it >= 2:
                for nme, ms in ctx.stage_ms().items():
                    acc.setdefault(nme, []).append(ms)
        ctr = ctx.counters()
    for k, v in sets:
        del os.environ[k]
    print(json.dumps({"arm": arm, "workload": a.workload, "stage_ms_median": {k: round(statistics.median(v), 4) for k, v in acc.items()},
                      "n_kept": ctr["n_kept"], "n_edges": ctr["n_edges"], "n_sweeps": ctr["n_sweeps"], "n_block_pairs": ctr["n_block_pairs"]}))
