"""GPU parity of the device-side host feed (umigpu_push_bam_records): raw BAM records in, kept record numbers out,
against the oracle's record decode + dedup; plus the file-level driver BAM -> BAM."""
import random

import numpy as np
import pytest

import oracle_lib as O
import umigpu
from bam_fixtures import make_bam
from umigpu import bamio

pytestmark = pytest.mark.gpu


def oracle_from_records(recs, umi_len, sep, use_mapq, algo, merge, k, p):
    idx, tid, pos, rev, umi, score = [], [], [], [], [], []
    for i, r in enumerate(recs):
        d = O.bam_decode(r, umi_len, sep, use_mapq)
        if d["valid"]:
            idx.append(i); tid.append(d["tid"]); pos.append(d["pos"]); rev.append(d["rev"]); umi.append(d["umi"]); score.append(d["score"])
    a = np.frombuffer(b"".join(umi), np.uint8).reshape(len(umi), umi_len)
    kept, _, ctr = O.dedup(tid, pos, rev, a, score, algo, merge, k, p)
    return [idx[j] for j in kept.tolist()], ctr, len(recs) - len(idx)


@pytest.mark.parametrize("merge,alphabet,umi_len", [(umigpu.MERGE_AVGQUAL, "ACGT", 8), (umigpu.MERGE_MAPQUAL, "ACGTN", 6), (umigpu.MERGE_ANY, "ACG", 10)])
def test_push_bam_records_matches_oracle(merge, alphabet, umi_len):
    rng = random.Random(umi_len)
    header, recs, _ = make_bam(rng, 6000, umi_len=umi_len, alphabet=alphabet)
    buf = header + b"".join(recs)
    offs, _ = bamio.record_offsets(buf, len(header))
    for chunk in (0, 1700):
        with umigpu.Context(umi_len, 1, 0.5, umigpu.ALGO_DIR, merge) as ctx:
            nun = 0
            if chunk:
                for s in range(0, len(recs), chunk):
                    e = min(len(recs), s + chunk)
                    nun += bamio.push_bam(ctx, buf, offs[s: e + 1], ord("_"), s)
            else:
                nun = bamio.push_bam(ctx, buf, offs, ord("_"), 0)
            kept, _, ctr = ctx.finish()
        okept, octr, ounmapped = oracle_from_records(recs, umi_len, ord("_"), merge == umigpu.MERGE_MAPQUAL, O.ALGO_DIR, merge, 1, 0.5)
        assert kept.astype(np.int64).tolist() == okept
        assert nun == ounmapped == ctr["n_unmapped"] and ctr["total_reads"] == len(recs)
        assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"] and ctr["max_umis"] == octr["max_umis"]


def test_bam_feed_errors_match_reference_panics():
    hdr = bamio.make_header(["c"], [1000])
    def run(qname, umi_len=4):
        rec = bamio.make_record(0, 10, 0, 30, qname, [(0, 10)], 10, bytes(10))
        buf = hdr + rec
        offs, _ = bamio.record_offsets(buf, len(hdr))
        with umigpu.Context(umi_len) as ctx:
            bamio.push_bam(ctx, buf, offs)
    with pytest.raises(umigpu.UmiGpuError, match="failed to get the umi"):
        run(b"noseparator")
    with pytest.raises(umigpu.UmiGpuError, match="too short"):
        run(b"r_AC")
    with pytest.raises(umigpu.UmiGpuError, match="Unknown character"):
        run(b"r_ACGx")
    run(b"r_ACGT")      # fine


def test_mixed_ascii_and_bam_chunks():
    rng = random.Random(9)
    header, recs, _ = make_bam(rng, 2000, umi_len=8, unmapped_rate=0.1)
    buf = header + b"".join(recs)
    offs, _ = bamio.record_offsets(buf, len(header))
    h = 900
    # first 900 records decoded by the oracle and pushed as SoA, the rest as raw records
    idx, tid, pos, rev, umi, score = [], [], [], [], [], []
    for i, r in enumerate(recs[:h]):
        d = O.bam_decode(r, 8, ord("_"), False)
        if d["valid"]:
            idx.append(i); tid.append(d["tid"]); pos.append(d["pos"]); rev.append(d["rev"]); umi.append(d["umi"]); score.append(d["score"])
    with umigpu.Context(8) as ctx:
        a = np.frombuffer(b"".join(umi), np.uint8).reshape(len(umi), 8)
        ctx.push_reads(np.array(tid, np.int32), np.array(pos, np.int64), np.array(rev, np.uint8), a, np.array(score, np.int32), None, 0)
        bamio.push_bam(ctx, buf, offs[h:], ord("_"), 5000)
        kept, _, _ = ctx.finish()
    okept, _, _ = oracle_from_records(recs, 8, ord("_"), False, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    pos_of = {r: j for j, r in enumerate(idx)}
    expect = [pos_of[r] if r < h else 5000 + (r - h) for r in okept]
    assert kept.astype(np.int64).tolist() == expect


@pytest.mark.parametrize("keep_unmapped", [False, True])
def test_file_level_bam_to_bam(tmp_path, keep_unmapped):
    rng = random.Random(12)
    header, recs, truth = make_bam(rng, 4000, umi_len=8)
    inp, out = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    bamio.bgzf_write_all(inp, header + b"".join(recs))
    args = umigpu.Cli(input=inp, output=out, k=1, algo_str="dir", merge_str="avgqual", data_str="naive", keep_unmapped=keep_unmapped)
    ctr = bamio.deduplicate_and_merge(args)
    back = bamio.bgzf_read_all(out)
    hdr, _, first = bamio.parse_header(back)
    assert hdr == header
    offs, _ = bamio.record_offsets(back, first)
    got = [bytes(back[int(offs[i]): int(offs[i + 1])]) for i in range(len(offs) - 1)]
    okept, octr, _ = oracle_from_records(recs, 8, ord("_"), False, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    keep = set(okept) | ({i for i, t in enumerate(truth) if t["unmapped"]} if keep_unmapped else set())
    assert got == [recs[i] for i in sorted(keep)]              # byte-identical records, input order
    assert ctr["n_kept"] == len(okept) == octr["n_kept"]


def _cli(*argv):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "umi-collapse-rs_b200", "host", "umicollapse_gpu")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    return subprocess.run([exe, *argv], capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("algo,merge", [("dir", "avgqual"), ("adj", "mapqual"), ("cc", "any")])
def test_cpp_cli_bam_matches_oracle(tmp_path, algo, merge):
    """The compiled C++ twin of the reference CLI: same flags as the README example of the reference
    (--mode bam --data naive --merge avgqual --num-threads N), records byte-identical, input order."""
    rng = random.Random(21)
    header, recs, truth = make_bam(rng, 5000, umi_len=8)
    inp, out = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    bamio.bgzf_write_all(inp, header + b"".join(recs))
    r = _cli("--mode", "bam", "-i", inp, "-o", out, "--data", "naive", "--algo", algo, "--merge", merge, "-k", "1", "--num-threads", "4")
    assert r.returncode == 0, r.stderr
    back = bamio.bgzf_read_all(out)
    hdr, _, first = bamio.parse_header(back)
    offs, _ = bamio.record_offsets(back, first)
    got = [bytes(back[int(offs[i]): int(offs[i + 1])]) for i in range(len(offs) - 1)]
    oalgo = {"dir": O.ALGO_DIR, "adj": O.ALGO_ADJ_REF, "cc": O.ALGO_CC}[algo]
    omerge = {"any": O.MERGE_ANY, "avgqual": O.MERGE_AVGQUAL, "mapqual": O.MERGE_MAPQUAL}[merge]
    okept, octr, ounm = oracle_from_records(recs, 8, ord("_"), merge == "mapqual", oalgo, omerge, 1, 0.5)
    assert hdr == header and got == [recs[i] for i in okept]
    # the reference's end-of-run counters (deduplicate_sam.rs:243-267)
    assert f"Number of input reads: {len(recs)}" in r.stderr
    assert f"Number of removed unmapped reads: {ounm}" in r.stderr
    assert f"Number of unique alignment positions: {octr['n_buckets']}" in r.stderr
    assert f"Number of UMIs: {octr['total_umis']}" in r.stderr
    assert f"Number of reads after deduplicating: {octr['n_kept']}" in r.stderr


def test_cpp_cli_fastq_single_bucket(tmp_path):
    """BASELINE config 4 shape: fastq, one global bucket (no reference behaviour: self-consistent with the oracle)."""
    rng = random.Random(5)
    n, L = 20000, 8
    umis = ["".join(rng.choice("ACGT") for _ in range(L)) for _ in range(n)]
    quals = ["".join(chr(33 + rng.randrange(2, 41)) for _ in range(30)) for _ in range(n)]
    recs = [f"@r{i}_{umis[i]} extra\n{'A' * 30}\n+\n{quals[i]}\n" for i in range(n)]
    inp, out = str(tmp_path / "in.fastq"), str(tmp_path / "out.fastq")
    open(inp, "w").write("".join(recs))
    r = _cli("--mode", "fastq", "-i", inp, "-o", out, "-k", "1", "--num-threads", "4")
    assert r.returncode == 0, r.stderr
    score = [sum(ord(c) - 33 for c in q) // len(q) for q in quals]
    a = np.frombuffer("".join(umis).encode(), np.uint8).reshape(n, L)
    okept, _, _ = O.dedup([0] * n, [0] * n, [0] * n, a, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5)
    assert open(out).read() == "".join(recs[i] for i in okept.tolist())


def test_cpp_cli_rejects_like_the_reference():
    r = _cli("-i", "a", "-o", "b", "--tag", "--two-pass")
    assert r.returncode != 0 and "Cannot track clusters with the two pass algorithm!" in r.stderr
    r = _cli("-i", "a", "-o", "b", "--paired", "--keep-unmapped")
    assert r.returncode != 0 and "Cannot keep unmapped reads with paired-end reads!" in r.stderr


def _parse_tags(rec: bytes):
    """(flag, {tag: value}) of one raw BAM record."""
    import struct
    l_name = rec[12]; n_cigar, flag = struct.unpack_from("<HH", rec, 16); l_seq, = struct.unpack_from("<i", rec, 20)
    p = 36 + l_name + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
    tags = {}
    while p < len(rec):
        tag, typ = rec[p:p + 2].decode(), chr(rec[p + 2]); p += 3
        if typ == "I":
            tags[tag] = struct.unpack_from("<I", rec, p)[0]; p += 4
        elif typ == "Z":
            e = rec.index(b"\0", p); tags[tag] = rec[p:e].decode(); p = e + 1
        else:
            raise AssertionError(typ)
    return flag, tags


def test_cpp_cli_tag_mode(tmp_path):
    """--tag as specified by the reference's help text (src/cli.rs:64-76; the reference itself writes nothing in this
    mode): every mapped read is written, non-consensus reads carry the duplicate flag, MI / RX on all, cs on the
    consensus read, su on the best read of each distinct UMI.  Expectations come from the oracle's cluster roots."""
    rng = random.Random(33)
    header, recs, truth = make_bam(rng, 3000, umi_len=8)
    inp, out = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    bamio.bgzf_write_all(inp, header + b"".join(recs))
    r = _cli("--mode", "bam", "-i", inp, "-o", out, "--algo", "dir", "--merge", "avgqual", "--tag", "--num-threads", "2")
    assert r.returncode == 0, r.stderr
    back = bamio.bgzf_read_all(out)
    _, _, first = bamio.parse_header(back)
    offs, _ = bamio.record_offsets(back, first)
    got = [bytes(back[int(offs[i]): int(offs[i + 1])]) for i in range(len(offs) - 1)]
    # oracle: decode, dedup with roots
    idx, tid, pos, rev, umi, score = [], [], [], [], [], []
    for i, rec in enumerate(recs):
        d = O.bam_decode(rec, 8, ord("_"), False)
        if d["valid"]:
            idx.append(i); tid.append(d["tid"]); pos.append(d["pos"]); rev.append(d["rev"]); umi.append(d["umi"]); score.append(d["score"])
    a = np.frombuffer(b"".join(umi), np.uint8).reshape(len(umi), 8)
    kept, roots, _ = O.dedup(tid, pos, rev, a, score, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5, want_roots=True)
    assert len(got) == len(idx)                                        # every mapped read is written
    roots = roots.tolist()
    csize = {}
    for rt in roots:
        csize[rt] = csize.get(rt, 0) + 1
    groups = {}
    for j in range(len(idx)):
        groups.setdefault((tid[j], pos[j], rev[j], umi[j]), []).append(j)
    urep, same = {}, {}
    for key, members in groups.items():
        best = min(m for m in members if score[m] == max(score[x] for x in members))
        urep[best] = True; same[best] = len(members)
    mi_of_root = {}
    for j, rec in enumerate(got):
        flag, tags = _parse_tags(rec)
        orig = recs[idx[j]]
        assert rec[4:36 + orig[12]] [:14] == orig[4:18]               # untouched fixed fields before the flag
        is_root = roots[j] == j
        assert bool(flag & 0x400) == (not is_root)
        assert tags["RX"] == umi[roots[j]].decode()
        mi_of_root.setdefault(roots[j], tags["MI"])
        assert tags["MI"] == mi_of_root[roots[j]]
        assert ("cs" in tags) == is_root and (not is_root or tags["cs"] == csize[j])
        assert ("su" in tags) == (j in urep) and (j not in urep or tags["su"] == same[j])
    assert len(set(mi_of_root.values())) == len(mi_of_root) == len(kept)      # one id per cluster


@pytest.mark.parametrize("batch_blocks", ["1", "3", "1024"])
def test_cpp_cli_two_pass_streaming_equals_in_memory(tmp_path, batch_blocks):
    """--two-pass (src/cli.rs:45-48): the input is streamed twice with bounded host memory; the output must carry
    exactly the records of the in-memory mode.  Small batches force records to straddle batch boundaries."""
    import os
    rng = random.Random(44)
    header, recs, truth = make_bam(rng, 6000, umi_len=8)
    inp = str(tmp_path / "in.bam")
    # small BGZF blocks so that many blocks / batches exist
    bamio.bgzf_write_all(inp, header + b"".join(recs), block=2000)
    outs = []
    for extra in ([], ["--two-pass"]):
        out = str(tmp_path / f"out{len(extra)}.bam")
        env_backup = os.environ.get("UMICOLLAPSE_BATCH_BLOCKS")
        os.environ["UMICOLLAPSE_BATCH_BLOCKS"] = batch_blocks
        try:
            r = _cli("--mode", "bam", "-i", inp, "-o", out, "--algo", "dir", "--merge", "avgqual", "--keep-unmapped", "--num-threads", "3", *extra)
        finally:
            if env_backup is None:
                os.environ.pop("UMICOLLAPSE_BATCH_BLOCKS", None)
            else:
                os.environ["UMICOLLAPSE_BATCH_BLOCKS"] = env_backup
        assert r.returncode == 0, r.stderr
        outs.append((bamio.bgzf_read_all(out), r.stderr))
    assert outs[0][0] == outs[1][0]
    for line in ("Number of input reads", "Number of removed unmapped reads", "Number of reads after deduplicating"):
        a = [l for l in outs[0][1].splitlines() if line in l]
        b = [l for l in outs[1][1].splitlines() if line in l]
        assert a == b and a
