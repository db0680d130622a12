"""End-to-end measurement of the compiled CLI twin on a synthetic BAM ("also reported end-to-end incl. BAM I/O").
Builds a coordinate-sorted BAM from the C2 generator (scaled), runs umicollapse_gpu on it and prints one JSON line with
the wall time, reads/s and the CLI's own phase times.  usage: bench_cli.py [scale=0.1] [threads=16]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "umi-collapse-rs_b200"))
from umigpu import bamio, synth  # noqa: E402


def build_bam(path, scale):
    d, cfg = synth.generate_config("C2", device="cpu", scale=scale)
    tid = d["tid"].numpy().astype(np.int32); pos = d["pos"].numpy().astype(np.int64); rev = d["rev"].numpy().astype(np.uint8)
    umi = d["umi"].numpy(); score = d["score"].numpy().astype(np.uint8)
    n, L = umi.shape
    read_len = 100
    name_len = 1 + 9 + 1 + L + 1                      # r#########_UMI\0
    rec_len = 4 + 32 + name_len + 4 + (read_len + 1) // 2 + read_len
    rec = np.zeros((n, rec_len), np.uint8)
    def put(off, arr, dt):
        rec[:, off: off + np.dtype(dt).itemsize] = np.ascontiguousarray(arr, dt).view(np.uint8).reshape(n, -1)
    put(0, np.full(n, rec_len - 4), "<i4")
    put(4, tid, "<i4")
    # the generator's position is the unclipped 5' position: reverse reads end there (100M: pos + 99)
    put(8, np.where(rev == 1, pos - (read_len - 1), pos), "<i4")
    rec[:, 12] = name_len; rec[:, 13] = 30
    put(14, np.full(n, 4680), "<u2"); put(16, np.full(n, 1), "<u2"); put(18, np.where(rev == 1, 16, 0), "<u2")
    put(20, np.full(n, read_len), "<i4"); put(24, np.full(n, -1), "<i4"); put(28, np.full(n, -1), "<i4")
    rec[:, 36] = ord("r")
    idx = np.arange(n)
    for k in range(9):
        rec[:, 37 + k] = ord("0") + (idx // 10 ** (8 - k)) % 10
    rec[:, 46] = ord("_")
    rec[:, 47: 47 + L] = umi
    co = 36 + name_len
    put(co, np.full(n, (read_len << 4) | 0), "<u4")
    rng = np.random.default_rng(1)
    rec[:, co + 4: co + 4 + (read_len + 1) // 2] = rng.integers(0x11, 0x89, (n, (read_len + 1) // 2), dtype=np.uint8)       # random bases
    rec[:, co + 4 + (read_len + 1) // 2:] = np.clip(score[:, None].astype(np.int16) + rng.integers(-6, 7, (n, read_len), dtype=np.int16), 2, 41).astype(np.uint8)
    header = bamio.make_header(["chr1"], [1 << 29])
    t0 = time.time()
    bamio.bgzf_write_all(path, header + rec.tobytes())
    return n, rec_len, time.time() - t0


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    threads = sys.argv[2] if len(sys.argv) > 2 else str(min(os.cpu_count() or 1, 16))
    exe = os.environ.get("UMICOLLAPSE_EXE") or os.path.join(REPO, "umi-collapse-rs_b200", "host", "umicollapse_gpu")
    with tempfile.TemporaryDirectory() as td:
        inp, out = os.path.join(td, "in.bam"), os.path.join(td, "out.bam")
        n, rec_len, t_build = build_bam(inp, scale)
        runs = []
        for mode in ([], ["--two-pass"]) if not os.environ.get("UMICOLLAPSE_EXE_B") else ([], ["__B__"]):
            for rep in range(3 if os.environ.get("UMICOLLAPSE_EXE_B") else 2):          # first run warms the page cache and the CUDA driver
                t0 = time.time()
                this_exe = os.environ["UMICOLLAPSE_EXE_B"] if mode == ["__B__"] else exe
                r = subprocess.run([this_exe, "--mode", "bam", "-i", inp, "-o", out, "--algo", "dir", "--merge", "avgqual", "-k", "1",
                                    "--num-threads", threads, *([] if mode == ["__B__"] else mode)], capture_output=True, text=True)
                dt = time.time() - t0
                assert r.returncode == 0, r.stderr
            phases = {m.group(1): float(m.group(2)) for m in re.finditer(r"phase (.+?): ([0-9.]+) s", r.stderr)}
            kept = int(re.search(r"Number of reads after deduplicating: (\d+)", r.stderr).group(1))
            runs.append({"mode": ("A/B: " + os.path.basename(this_exe)) if os.environ.get("UMICOLLAPSE_EXE_B") else ("two-pass (streaming)" if mode else "in-memory"), "wall_s": dt, "reads_per_s": n / dt, "phases_s": phases, "kept": kept})
        print(json.dumps({"tool": "umicollapse_gpu (C++ CLI twin)", "workload": f"C2 generator x{scale}: {n} reads, {rec_len} B records, 100M CIGAR, coordinate-sorted BAM",
                          "input_bytes": os.path.getsize(inp), "output_bytes": os.path.getsize(out), "threads": int(threads),
                          "host_cores": os.cpu_count(), "runs": runs}))


if __name__ == "__main__":
    main()
