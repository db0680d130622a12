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


# ---- paired-end mode: PairedAlignment key + the filters of deduplicate_sam.rs:96-129 ----
def oracle_paired_from_records(recs, umi_len, sep, use_mapq, algo, merge, k, p, remove_unpaired, remove_chimeric):
    idx, tid, pos, rev, tlen, umi, score = [], [], [], [], [], [], []
    n = dict(mates=0, unmapped=0, unpaired=0, chimeric=0)
    for i, r in enumerate(recs):
        d = O.bam_decode(r, umi_len, sep, use_mapq, True, remove_unpaired, remove_chimeric)
        n["mates"] += bool(d["cls"] & O.CLS_MATE); n["unmapped"] += bool(d["cls"] & O.CLS_UNMAPPED)
        n["unpaired"] += bool(d["cls"] & O.CLS_UNPAIRED); n["chimeric"] += bool(d["cls"] & O.CLS_CHIMERIC)
        if d["valid"]:
            idx.append(i); tid.append(d["tid"]); pos.append(d["pos"]); rev.append(d["rev"]); tlen.append(d["tlen"])
            umi.append(d["umi"]); score.append(d["score"])
    a = np.frombuffer(b"".join(umi), np.uint8).reshape(len(umi), umi_len)
    kept, _, ctr = O.dedup(tid, pos, rev, a, score, algo, merge, k, p, tlen=tlen)
    return [idx[j] for j in kept.tolist()], ctr, n


def test_push_reads_paired_matches_oracle():
    rng = np.random.default_rng(5)
    n, L = 20000, 8
    tid = rng.integers(0, 3, n).astype(np.int32); pos = rng.integers(0, 40, n).astype(np.int64) * 7 - 50
    rev = rng.integers(0, 2, n).astype(np.uint8); tlen = rng.choice(np.array([-300, -1, 0, 150, 151, 2 ** 31 - 1], np.int64), n)
    fam = rng.integers(0, 4, (60, L))
    umi = np.frombuffer(b"ACGT", np.uint8)[np.where(rng.random((n, L)) < 0.05, rng.integers(0, 4, (n, L)), fam[rng.integers(0, 60, n)])]
    score = rng.integers(0, 41, n).astype(np.int32)
    for algo in (umigpu.ALGO_DIR, umigpu.ALGO_ADJ_UPSTREAM, umigpu.ALGO_CC):
        with umigpu.Context(L, 1, 0.5, algo, umigpu.MERGE_AVGQUAL) as ctx:
            half = n // 2
            ctx.push_reads(tid[:half], pos[:half], rev[:half], umi[:half], score[:half], tlen=tlen[:half])
            ctx.push_reads(tid[half:], pos[half:], rev[half:], umi[half:], score[half:], first_read_index=half, tlen=tlen[half:])
            kept, _, ctr = ctx.finish()
        okept, _, octr = O.dedup(tid, pos, rev, umi, score, algo, umigpu.MERGE_AVGQUAL, 1, 0.5, tlen=tlen)
        assert kept.astype(np.int64).tolist() == okept.tolist()
        assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"]
        # and the template length really splits buckets
        _, _, uctr = O.dedup(tid, pos, rev, umi, score, algo, umigpu.MERGE_AVGQUAL, 1, 0.5)
        assert uctr["n_buckets"] < octr["n_buckets"]
    with umigpu.Context(L) as ctx:          # paired and unpaired chunks cannot share a batch
        ctx.push_reads(tid[:10], pos[:10], rev[:10], umi[:10], score[:10], tlen=tlen[:10])
        with pytest.raises(umigpu.UmiGpuError, match="cannot be mixed"):
            ctx.push_reads(tid[10:20], pos[10:20], rev[10:20], umi[10:20], score[10:20], first_read_index=10)


@pytest.mark.parametrize("remove_unpaired,remove_chimeric", [(False, False), (True, False), (False, True), (True, True)])
def test_push_bam_records_paired_filters(remove_unpaired, remove_chimeric):
    from bam_fixtures import make_paired_bam
    rng = random.Random(11 + remove_unpaired + 2 * remove_chimeric)
    header, recs = make_paired_bam(rng, 4000)
    buf = header + b"".join(recs)
    offs, _ = bamio.record_offsets(buf, len(header))
    fl = umigpu.FLAG_PAIRED | (umigpu.FLAG_REMOVE_UNPAIRED if remove_unpaired else 0) | (umigpu.FLAG_REMOVE_CHIMERIC if remove_chimeric else 0)
    okept, octr, on = oracle_paired_from_records(recs, 8, ord("_"), False, O.ALGO_DIR, umigpu.MERGE_AVGQUAL, 1, 0.5, remove_unpaired, remove_chimeric)
    assert on["mates"] and on["unmapped"] and on["unpaired"] and on["chimeric"]
    for chunk in (0, 1111):
        with umigpu.Context(8, 1, 0.5, umigpu.ALGO_DIR, umigpu.MERGE_AVGQUAL, flags=fl) as ctx:
            nun = 0
            for s in range(0, len(recs), chunk or len(recs)):
                e = min(len(recs), s + (chunk or len(recs)))
                nun += bamio.push_bam(ctx, buf, offs[s: e + 1], ord("_"), s)
            kept, _, ctr = ctx.finish()
        assert kept.astype(np.int64).tolist() == okept
        assert nun == on["unmapped"] == ctr["n_unmapped"]
        assert ctr["total_reads"] == len(recs) - on["mates"] and ctr["n_mates_skipped"] == on["mates"]
        assert ctr["n_unpaired"] == on["unpaired"] and ctr["n_chimeric"] == on["chimeric"]
        assert ctr["n_buckets"] == octr["n_buckets"] and ctr["total_umis"] == octr["total_umis"] and ctr["max_umis"] == octr["max_umis"]


def _expected_paired_output(recs, okept):
    """UcWriter::write + write_reversed (deduplicate_sam.rs:382-462) restated on the test side: the kept reads plus, for
    every kept paired read, the first mapped last-in-template record with the same name at (mate ref, mate pos)."""
    import struct
    def f(rec):
        flag, = struct.unpack_from("<H", rec, 18)
        tid, pos = struct.unpack_from("<ii", rec, 4); mtid, mpos = struct.unpack_from("<ii", rec, 24)
        return rec[36: 36 + rec[12] - 1], flag, tid, pos, mtid, mpos
    want = set()
    for i in okept:
        name, flag, _, _, mtid, mpos = f(recs[i])
        if flag & 1:
            want.add((name, mtid, mpos))
    keep = set(okept)
    for i, rec in enumerate(recs):
        name, flag, tid, pos, _, _ = f(rec)
        if not (flag & 4) and (flag & 1) and (flag & 0x80) and not (flag & 8) and (name, tid, pos) in want:
            want.remove((name, tid, pos))
            keep.add(i)
    return [recs[i] for i in sorted(keep)]


@pytest.mark.parametrize("extra", [[], ["--remove-unpaired"], ["--remove-chimeric", "--two-pass"], ["--two-pass"]])
def test_cpp_cli_paired(tmp_path, extra):
    """--paired (src/cli.rs:49-60): first reads are deduplicated with the template length in the key, their mates follow."""
    import os
    from bam_fixtures import make_paired_bam
    rng = random.Random(77)
    header, recs = make_paired_bam(rng, 5000)
    inp, out = str(tmp_path / "in.bam"), str(tmp_path / "out.bam")
    bamio.bgzf_write_all(inp, header + b"".join(recs), block=3000)
    os.environ["UMICOLLAPSE_BATCH_BLOCKS"] = "5"
    try:
        r = _cli("--mode", "bam", "-i", inp, "-o", out, "--algo", "dir", "--merge", "avgqual", "--paired", "--num-threads", "3", *extra)
    finally:
        os.environ.pop("UMICOLLAPSE_BATCH_BLOCKS", None)
    assert r.returncode == 0, r.stderr
    ru, rc = "--remove-unpaired" in extra, "--remove-chimeric" in extra
    okept, octr, on = oracle_paired_from_records(recs, 8, ord("_"), False, O.ALGO_DIR, O.MERGE_AVGQUAL, 1, 0.5, ru, rc)
    back = bamio.bgzf_read_all(out)
    hdr, _, first = bamio.parse_header(back)
    offs, _ = bamio.record_offsets(back, first)
    got = [bytes(back[int(offs[i]): int(offs[i + 1])]) for i in range(len(offs) - 1)]
    expect = _expected_paired_output(recs, okept)
    assert len(expect) > len(okept)                      # mates were added
    assert hdr == header and got == expect
    assert f"Number of input reads: {len(recs) - on['mates']}" in r.stderr
    assert f"Number of removed unmapped reads: {on['unmapped']}" in r.stderr
    assert f"Number of unpaired reads: {on['unpaired']}" in r.stderr
    assert f"Number of chimeric reads: {on['chimeric']}" in r.stderr
    assert f"Number of unique alignment positions: {octr['n_buckets']}" in r.stderr
    assert f"Number of reads after deduplicating: {octr['n_kept']}" in r.stderr
    # the Python driver writes the same file
    out2 = str(tmp_path / "out2.bam")
    args = umigpu.Cli(input=inp, output=out2, k=1, algo_str="dir", merge_str="avgqual", data_str="naive", paired=True,
                      remove_unpaired=ru, remove_chimeric=rc)
    bamio.deduplicate_and_merge(args)
    assert bamio.bgzf_read_all(out2) == back

