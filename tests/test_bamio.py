"""CPU tests of the host feed: BGZF/BAM byte handling (pure host code), the record walker exported by the C ABI
(a host function, no GPU), and the oracle's record decode against the literal Python restatement."""
import random

import numpy as np

import oracle_lib as O
import ref_literal as R
from bam_fixtures import make_bam
from umigpu import bamio


def test_bgzf_roundtrip_and_header(tmp_path):
    rng = random.Random(1)
    header, recs, _ = make_bam(rng, 500)
    data = header + b"".join(recs)
    p = str(tmp_path / "t.bam")
    bamio.bgzf_write_all(p, data, block=4096)
    back = bamio.bgzf_read_all(p)
    assert back == data
    hdr, names, first = bamio.parse_header(back)
    assert hdr == header and names == ["chr0", "chr1", "chr2"] and first == len(header)


def test_record_offsets_walker():
    rng = random.Random(2)
    header, recs, _ = make_bam(rng, 300)
    buf = header + b"".join(recs)
    offs, consumed = bamio.record_offsets(buf, len(header))
    assert len(offs) == 301 and consumed == len(buf) - len(header)
    exp = np.cumsum([len(header)] + [len(r) for r in recs])
    assert offs.tolist() == exp.tolist()
    # a trailing partial record is left for the next call
    offs2, consumed2 = bamio.record_offsets(buf[:-5], len(header))
    assert len(offs2) == 300 and consumed2 == len(buf) - len(header) - len(recs[-1])


def test_oracle_record_decode_matches_literal():
    rng = random.Random(3)
    _, recs, truth = make_bam(rng, 400, umi_len=7, alphabet="ACGTN")
    for rec, t in zip(recs, truth):
        for use_mapq in (False, True):
            d = O.bam_decode(rec, 7, ord("_"), use_mapq)
            assert d["valid"] == (0 if t["unmapped"] else 1)
            if t["unmapped"]:
                continue
            assert d["tid"] == t["tid"] and d["rev"] == int(t["rev"]) and d["umi"] == t["umi"].encode()
            assert d["pos"] == R.unclipped_pos(t["pos"], t["rev"], t["cigar"])
            assert d["score"] == (t["mapq"] if use_mapq else R.avg_qual(t["qual"]))


def test_umi_length_autodetect():
    rng = random.Random(4)
    header, recs, _ = make_bam(rng, 20, umi_len=9, unmapped_rate=0.5)
    buf = header + b"".join(recs)
    offs, _ = bamio.record_offsets(buf, len(header))
    assert bamio.autodetect_umi_length(buf, offs, ord("_")) == 9


def test_umi_length_autodetect_follows_the_reference_regex():
    """^(?:.*?)_([ATCGN]+)(?:.*?)$ caseless: the first separator FOLLOWED BY A LETTER decides the length, even though
    get_umi later cuts after the first separator (utils/read.rs:67 vs :100-101)."""
    hdr = bamio.make_header(["c"], [1000])
    def length(qname):
        rec = bamio.make_record(0, 10, 0, 30, qname, [(0, 10)], 10, bytes(10))
        buf = hdr + rec
        offs, _ = bamio.record_offsets(buf, len(hdr))
        return bamio.autodetect_umi_length(buf, offs, ord("_"))
    assert length(b"read_ACGTAC") == 6
    assert length(b"read_ACGTAC_tail") == 6
    assert length(b"read_1_ACGT") == 4            # first separator is followed by a digit: the regex moves on
    assert length(b"read_acgtn") == 5             # caseless
    import pytest
    with pytest.raises(ValueError):
        length(b"read_12")


def test_oracle_paired_filters_match_literal():
    """deduplicate_sam.rs:96-129: the C oracle's record classes against the statement-by-statement restatement."""
    import struct
    from bam_fixtures import make_paired_bam
    rng = random.Random(8)
    _, recs = make_paired_bam(rng, 1500)
    seen = set()
    for ru in (False, True):
        for rc in (False, True):
            for rec in recs:
                tid, = struct.unpack_from("<i", rec, 4); flag, = struct.unpack_from("<H", rec, 18)
                mtid, = struct.unpack_from("<i", rec, 24); tlen, = struct.unpack_from("<i", rec, 32)
                ok, c = R.paired_filter(flag, tid, mtid, ru, rc)
                d = O.bam_decode(rec, 8, ord("_"), False, True, ru, rc)
                assert d["valid"] == int(ok)
                assert bool(d["cls"] & O.CLS_MATE) == (c["total"] == 0)
                assert bool(d["cls"] & O.CLS_UNMAPPED) == bool(c["unmapped"])
                assert bool(d["cls"] & O.CLS_UNPAIRED) == bool(c["unpaired"])
                assert bool(d["cls"] & O.CLS_CHIMERIC) == bool(c["chimeric"])
                if ok:
                    assert d["tlen"] == tlen
                seen.add((d["valid"], d["cls"]))
    assert len(seen) >= 6      # the fixture exercises every branch


def test_select_mates_follows_write_reversed():
    """UcWriter::write / write_reversed (deduplicate_sam.rs:382-462) on the host: for every kept paired read the first
    mapped last-in-template record with its name at (mate ref, mate pos) is written, once."""
    import struct
    from bam_fixtures import make_paired_bam
    rng = random.Random(19)
    header, recs = make_paired_bam(rng, 1200)
    buf = header + b"".join(recs)
    offs, _ = bamio.record_offsets(buf, len(header))
    def fields(i):
        rec = recs[i]
        flag, = struct.unpack_from("<H", rec, 18)
        tid, pos = struct.unpack_from("<ii", rec, 4); mtid, mpos = struct.unpack_from("<ii", rec, 24)
        return rec[36: 36 + rec[12] - 1], flag, tid, pos, mtid, mpos
    # "kept" = every record that passes the paired filters (what reaches the writer when nothing is a duplicate)
    kept = [i for i in range(len(recs)) if R.paired_filter(fields(i)[1], fields(i)[2], fields(i)[4], False, False)[0]]
    mates = bamio.select_mates(buf, offs, np.array(kept, np.int64))
    assert mates == sorted(set(mates)) and not set(mates) & set(kept)
    want = {(fields(i)[0], fields(i)[4], fields(i)[5]) for i in kept if fields(i)[1] & 1}
    seen = set()
    for m in mates:
        name, flag, tid, pos, _, _ = fields(m)
        assert (flag & 1) and (flag & 0x80) and not (flag & 4) and not (flag & 8)      # :429-433
        assert (name, tid, pos) in want and (name, tid, pos) not in seen                 # :444-458: written once
        seen.add((name, tid, pos))
    # nothing that could have been written was left out, and a repeated mate record yields its first copy
    for i in range(len(recs)):
        name, flag, tid, pos, _, _ = fields(i)
        if (flag & 1) and (flag & 0x80) and not (flag & 4) and not (flag & 8) and (name, tid, pos) in want:
            assert (name, tid, pos) in seen
            first_copy = min(j for j in range(len(recs)) if recs[j] == recs[i])
            assert first_copy in mates
    assert len(mates) > 500
