"""The small cases run under compute-sanitizer (memcheck / racecheck / initcheck): every kernel family of the path at sizes
the tools finish in a minute — segmented sort (windows + big segments), generic sort, multi-index neighbour passes,
plain / frontier / two-phase clustering, and a 2-rank shard group on one device with the hot bucket split."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("umi-collapse-rs_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import oracle_lib as O          # noqa: E402
import umigpu                   # noqa: E402
from umigpu import synth        # noqa: E402


def case(name, scale, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    d, cfg = synth.generate_config(name, device="cpu", scale=scale, **kw)
    h = {k: v.numpy() for k, v in d.items()}
    with umigpu.Context(cfg["umi_len"], cfg["k"], 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, 0) as ctx:
        ctx.push_reads(h["tid"], h["pos"], h["rev"], h["umi"], h["score"])
        kept, _, ctr = ctx.finish()
    okept, _, _ = O.dedup(h["tid"], h["pos"], h["rev"], h["umi"], h["score"], O.ALGO_DIR, O.MERGE_AVGQUAL, cfg["k"], 0.5)
    assert kept.astype(np.int64).tolist() == okept.tolist(), name
    for k in (env or {}):
        del os.environ[k]
    print(name, scale, env, "ok:", ctr["total_reads"], "reads", ctr["n_edges"], "edges", ctr["n_sweeps"], "sweeps", flush=True)
    return h, cfg, okept


case("C1", 0.02)
h, cfg, okept = case("C2", 0.002)
case("C2", 0.002, {"UMIGPU_NO_SEG_SORT": "1"})
case("C2", 0.002, {"UMIGPU_SV_MIN_EDGES": "0", "UMIGPU_PLAIN_ROUNDS": "0"})
case("C2", 0.002, {"UMIGPU_FRONTIER_FORCE": "1", "UMIGPU_FRONTIER_MIN_EDGES": "1"})
case("C4", 0.001)
os.environ["UMIGPU_HOT_MIN_READS"] = "2000"
with umigpu.Group(cfg["umi_len"], [0, 0]) as g:
    kept, ctr, _ = g.dedup(h["tid"], h["pos"], h["rev"], h["umi"], h["score"])
assert kept.astype(np.int64).tolist() == okept.tolist()
print("group [0,0] with hot split ok", flush=True)
