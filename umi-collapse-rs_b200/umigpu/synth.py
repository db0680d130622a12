"""Seeded synthetic inputs of the shapes named in BASELINE.json.

Everything is generated with torch ops so that the 50M / 200M-read configs can be produced on the
GPU in seconds; with device="cpu" the same code yields the small parity-test inputs.  Output is the
SoA the hot path consumes (what HOT LOOP A of src/deduplicate_sam.rs:93-177 extracts per record):
tid int32, unclipped pos int64, strand uint8, UMI ASCII uint8[n, L], score int32.
"""
from __future__ import annotations

import torch

_ACGT = (65, 67, 71, 84)

CONFIGS = {
    # name: reads, umi_len, loci, zipf_s, family_mean, err, k, algo, n_contigs, single_bucket
    "C1": dict(n_reads=1_000_000, umi_len=10, n_loci=5_000, zipf_s=0.0, family=4.0, err=0.01, k=1, algo="dir",
               both_strands=False, single_bucket=False),
    "C2": dict(n_reads=50_000_000, umi_len=12, n_loci=1_000_000, zipf_s=1.1, family=4.0, err=0.01, k=1, algo="dir",
               both_strands=True, single_bucket=False),
    "C3": dict(n_reads=50_000_000, umi_len=16, n_loci=1_000_000, zipf_s=0.8, family=4.0, err=0.02, k=2, algo="cc",
               both_strands=True, single_bucket=False),
    "C4": dict(n_reads=20_000_000, umi_len=12, n_loci=1, zipf_s=0.0, family=1.0, err=0.0, k=1, algo="dir",
               both_strands=False, single_bucket=True),
    "C5": dict(n_reads=200_000_000, umi_len=12, n_loci=4_000_000, zipf_s=1.1, family=4.0, err=0.01, k=1, algo="dir",
               both_strands=True, single_bucket=False),
}


def _mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrap-around arithmetic)."""
    x = (x ^ (x >> 30).bitwise_and(0x3FFFFFFFF)) * -4658895280553007687      # 0xBF58476D1CE4E5B9
    x = (x ^ (x >> 27).bitwise_and(0x1FFFFFFFFF)) * -7723592293110705685     # 0x94D049BB133111EB
    return x ^ (x >> 31).bitwise_and(0x1FFFFFFFF)


def generate(n_reads: int, umi_len: int, n_loci: int, zipf_s: float, family: float, err: float, seed: int,
             device: str = "cpu", both_strands: bool = True, single_bucket: bool = False, n_rate: float = 0.0,
             span: int = 1 << 28, n_contigs: int = 1, sort_by_coordinate: bool = True, **_unused):
    """Returns dict(tid, pos, rev, umi, score) of torch tensors on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = n_reads
    if single_bucket:
        locus = torch.zeros(n, dtype=torch.int64, device=dev)
        code = torch.randint(0, 4 ** umi_len, (n,), generator=g, device=dev, dtype=torch.int64)
        pos_of_locus = torch.zeros(1, dtype=torch.int64, device=dev)
        rev_of_locus = torch.zeros(1, dtype=torch.uint8, device=dev)
    else:
        k = torch.arange(1, n_loci + 1, device=dev, dtype=torch.float64)
        w = k.pow(-zipf_s) if zipf_s > 0 else torch.ones_like(k)
        cdf = torch.cumsum(w, 0)
        cdf = cdf / cdf[-1]
        u = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        locus = torch.searchsorted(cdf, u).clamp_(max=n_loci - 1)
        counts = torch.bincount(locus, minlength=n_loci)
        n_mol = torch.clamp((counts.to(torch.float64) / family).ceil().to(torch.int64), min=1)
        mol = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * n_mol[locus].to(torch.float64)).to(torch.int64)
        code = _mix64(locus * 0x632BE5AB + mol * 0x9E3779B1 + seed * 0x85EBCA6B) & (4 ** umi_len - 1)
        # loci at distinct sorted positions; hot (low-rank) loci are scattered over the span by a permutation
        pos_sorted = torch.sort(torch.randperm(span, generator=g, device=dev)[:n_loci] if span <= (1 << 24)
                                else torch.randint(0, span, (n_loci,), generator=g, device=dev, dtype=torch.int64)).values
        perm = torch.randperm(n_loci, generator=g, device=dev)
        pos_of_locus = pos_sorted[perm].to(torch.int64)
        rev_of_locus = (torch.randint(0, 2, (n_loci,), generator=g, device=dev, dtype=torch.int64) if both_strands
                        else torch.zeros(n_loci, dtype=torch.int64, device=dev)).to(torch.uint8)
    if err > 0:
        hit = torch.rand(n, generator=g, device=dev) < err
        where = torch.randint(0, umi_len, (n,), generator=g, device=dev, dtype=torch.int64)
        sub = torch.randint(1, 4, (n,), generator=g, device=dev, dtype=torch.int64)
        code = torch.where(hit, code ^ (sub << (2 * where)), code)
    pos = pos_of_locus[locus]
    rev = rev_of_locus[locus]
    contig_len = max(1, span // n_contigs)
    tid = (pos // contig_len).to(torch.int32) if n_contigs > 1 else torch.zeros(n, dtype=torch.int32, device=dev)
    if n_contigs > 1:
        pos = pos % contig_len
    score = torch.randint(2, 41, (n,), generator=g, device=dev, dtype=torch.int64).to(torch.int32)
    if sort_by_coordinate and not single_bucket:
        order = torch.argsort(tid.to(torch.int64) * (span * 2) + pos, stable=True)
        tid, pos, rev, code, score = tid[order], pos[order], rev[order], code[order], score[order]
    lut = torch.tensor(_ACGT, dtype=torch.uint8, device=dev)
    shifts = 2 * torch.arange(umi_len - 1, -1, -1, device=dev, dtype=torch.int64)
    umi = lut[((code.unsqueeze(1) >> shifts) & 3)]
    if n_rate > 0:
        nmask = torch.rand(n, umi_len, generator=g, device=dev) < n_rate
        umi = torch.where(nmask, torch.full_like(umi, 78), umi)
    return dict(tid=tid.contiguous(), pos=pos.contiguous(), rev=rev.contiguous(), umi=umi.contiguous(), score=score.contiguous())


def generate_config(name: str, seed: int | None = None, device: str = "cpu", scale: float = 1.0, **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    if scale != 1.0:
        cfg["n_reads"] = max(1000, int(cfg["n_reads"] * scale))
        if not cfg["single_bucket"]:
            cfg["n_loci"] = max(16, int(cfg["n_loci"] * scale))
    if seed is None:
        seed = int(name[1:])
    return generate(seed=seed, device=device, **cfg), cfg
