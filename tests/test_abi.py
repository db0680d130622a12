"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/umigpu.h declares, fails
loudly without a GPU, and the pure-host pieces (shard plan, CLI mirror) behave.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import umigpu
from umigpu import _lib as L

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "umigpu.h")).read()
    declared = sorted(set(re.findall(r"\b(umigpu_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(L.SYMBOLS)
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in umigpu.load().umigpu_version()


def test_rust_binding_declares_the_same_entry_points():
    """The FFI crate a reference maintainer would add (uncompiled here: no cargo) binds exactly the header's
    entry points — no stale or missing `extern "C"` item."""
    header = open(os.path.join(REPO, "include", "umigpu.h")).read()
    declared = sorted(set(re.findall(r"\b(umigpu_[a-z_0-9]+)\s*\(", header)))
    rust = open(os.path.join(REPO, "umi-collapse-rs_b200", "rust", "umigpu-sys", "src", "lib.rs")).read()
    bound = sorted(set(re.findall(r"\bfn\s+(umigpu_[a-z_0-9]+)\s*\(", rust)))
    assert bound == declared


def test_struct_layouts_match_header():
    assert C.sizeof(L.Config) == 40
    assert C.sizeof(L.Counters) == 136
    assert C.sizeof(L.Result) == 32 + 136 + 8


def test_sm100a_only_cubin():
    """The shipped library must contain sm_100a code and nothing else (no multi-arch dispatch)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(umigpu.UmiGpuError) as e:
        umigpu.Context(umi_len=12)
    assert e.value.code == L.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_config_before_touching_cuda():
    lib = umigpu.load()
    for kw in (dict(umi_len=0), dict(umi_len=33), dict(k=-1), dict(algo=9), dict(merge=7)):
        base = dict(k=1, percentage=0.5, algo=0, merge=1, umi_len=12, device=0, flags=0, reserved=0, stream=None)
        base.update(kw)
        cfg = L.Config(**base)
        h = C.c_void_p()
        rc = lib.umigpu_create(C.byref(cfg), C.byref(h))
        assert rc in (L.ERR_ARG, L.ERR_UNSUPPORTED) and not h.value


def test_cli_mirror_defaults_and_dispatch():
    """src/cli.rs defaults, main.rs:33-47 validation and :52-92 dispatch."""
    a = umigpu.Cli()
    assert (a.mode, a.k, a.percentage, a.algo_str, a.data_str, a.umi_separator) == ("bam", 1, 0.5, "dir", "ngrambktree", ord("_"))
    assert umigpu.resolve_cli(a) == (L.ALGO_DIR, L.MERGE_MAPQUAL)                 # bam -> mapqual
    assert umigpu.resolve_cli(umigpu.Cli(mode="fastq")) == (L.ALGO_DIR, L.MERGE_AVGQUAL)
    assert umigpu.resolve_cli(umigpu.Cli(algo_str="adj", merge_str="any", data_str="naive")) == (L.ALGO_ADJ, L.MERGE_ANY)
    with pytest.raises(ValueError):
        umigpu.resolve_cli(umigpu.Cli(algo_str="bogus"))
    with pytest.raises(ValueError):
        umigpu.resolve_cli(umigpu.Cli(track_clusters=True, two_pass=True))
    with pytest.raises(ValueError):
        umigpu.resolve_cli(umigpu.Cli(paired=True, keep_unmapped=True))


def test_shard_plan_keeps_buckets_whole_and_balances():
    rng = np.random.default_rng(0)
    n = 200_000
    locus = np.minimum((rng.pareto(1.1, n) * 3).astype(np.int64), 5000)
    tid = (locus % 3).astype(np.int32)
    pos = (locus * 13 - 100).astype(np.int64)
    rev = (locus % 2).astype(np.uint8)
    for shards in (1, 2, 4, 8):
        sh, cost = umigpu.shard_plan(tid, pos, rev, shards)
        assert sh.min() >= 0 and sh.max() < shards
        key = locus
        first = {}
        for kk, s in zip(key.tolist(), sh.tolist()):
            assert first.setdefault(kk, s) == s          # a bucket never spans shards
        counts = np.bincount(locus)
        c = counts.astype(np.int64) ** 2 + 64 * counts
        assert int(cost.sum()) == int(c.sum())
        # LPT bound: max load <= mean + largest item
        assert cost.max() <= cost.sum() / shards + c.max()


def test_cpp_cli_twin_validates_like_main_rs():
    """umi-collapse-rs_b200/host/umicollapse_gpu (no GPU needed for argument handling): required arguments,
    main.rs:41-47 panics, unknown flags."""
    import subprocess
    exe = os.path.join(REPO, "umi-collapse-rs_b200", "host", "umicollapse_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    def run(*argv):
        return subprocess.run([exe, *argv], capture_output=True, text=True, timeout=60)
    r = run("-i", "a")
    assert r.returncode != 0 and "required arguments" in r.stderr
    r = run("-i", "a", "-o", "b", "--tag", "--two-pass")
    assert r.returncode != 0 and "Cannot track clusters with the two pass algorithm!" in r.stderr
    r = run("-i", "a", "-o", "b", "--paired", "--keep-unmapped")
    assert r.returncode != 0 and "Cannot keep unmapped reads with paired-end reads!" in r.stderr
    r = run("-i", "a", "-o", "b", "--bogus")
    assert r.returncode != 0 and "unexpected argument" in r.stderr
    r = run("-i", "/nonexistent/in.bam", "-o", "/tmp/out.bam", "--data", "ngrambktree", "--num-threads", "2")
    assert r.returncode != 0 and "Invalid input path" in r.stderr          # deduplicate_sam.rs:78 expect("Invalid input path")
