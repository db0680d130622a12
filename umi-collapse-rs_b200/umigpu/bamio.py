"""Minimal BAM / BGZF reader and writer (zlib only — no htslib in this image) and the file-level driver that
mirrors DeduplicateInterface::deduplicate_and_merge (src/deduplicate_sam.rs:72-269) on top of the device feed
umigpu_push_bam_records.  Host code here only moves bytes: BGZF inflate/deflate, header parsing and writing the
surviving records; every per-record computation (unclipped position, UMI extraction, score, filter) and the
whole clustering run on the GPU."""
from __future__ import annotations

import ctypes as C
import gzip
import struct
import zlib

import numpy as np

from . import _lib as L
from .api import Cli, Context, resolve_cli

_EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_read_all(path: str) -> bytes:
    """BGZF is a series of gzip members; the gzip module reads them all."""
    with gzip.open(path, "rb") as f:
        return f.read()


def bgzf_write_all(path: str, data: bytes, level: int = 1, block: int = 0xff00):
    with open(path, "wb") as f:
        for s in range(0, len(data), block):
            chunk = data[s: s + block]
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            bsize = len(comp) + 25
            f.write(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, bsize))
            f.write(comp)
            f.write(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
        f.write(_EOF_BLOCK)


def parse_header(buf: bytes):
    """Returns (header_bytes, reference names, offset of the first alignment record)."""
    if buf[:4] != b"BAM\x01":
        raise ValueError("not a BAM stream")
    l_text, = struct.unpack_from("<i", buf, 4)
    off = 8 + l_text
    n_ref, = struct.unpack_from("<i", buf, off)
    off += 4
    names = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", buf, off)
        names.append(buf[off + 4: off + 4 + l_name - 1].decode())
        off += 4 + l_name + 4
    return buf[:off], names, off


def make_header(ref_names, ref_lens, text: str = "@HD\tVN:1.6\tSO:coordinate\n") -> bytes:
    out = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(ref_names))]
    for nme, ln in zip(ref_names, ref_lens):
        b = nme.encode() + b"\0"
        out += [struct.pack("<i", len(b)), b, struct.pack("<i", ln)]
    return b"".join(out)


def make_record(tid: int, pos: int, flag: int, mapq: int, qname: bytes, cigar, seq_len: int, qual: bytes,
                mtid: int = -1, mpos: int = -1, tlen: int = 0) -> bytes:
    """cigar = [(op, len)] with op codes M0 I1 D2 N3 S4 H5 P6 =7 X8; sequence bases are irrelevant to the path (all A)."""
    name = qname + b"\0"
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, 4680, len(cigar), flag, seq_len, mtid, mpos, tlen)
    body += name + b"".join(struct.pack("<I", (ln << 4) | op) for op, ln in cigar)
    body += bytes((seq_len + 1) // 2) + qual
    return struct.pack("<i", len(body)) + body


def record_offsets(buf, start: int = 0):
    """umigpu_bam_record_offsets over buf[start:]; offsets are relative to buf."""
    lib = L.load()
    arr = np.frombuffer(buf, dtype=np.uint8)
    n_max = max(1, (len(arr) - start) // 36 + 1)
    offs = np.zeros(n_max + 1, np.uint64)
    n, consumed = C.c_uint64(), C.c_uint64()
    L.check(lib.umigpu_bam_record_offsets(arr[start:].ctypes.data_as(C.c_void_p), len(arr) - start, offs.ctypes.data_as(C.c_void_p), n_max,
                                          C.byref(n), C.byref(consumed)))
    return offs[: n.value + 1] + np.uint64(start), int(consumed.value)


def push_bam(ctx: Context, buf, offsets: np.ndarray, umi_sep: int = ord("_"), first_read_index: int = 0) -> int:
    """umigpu_push_bam_records; returns the number of records dropped by the unmapped filter."""
    arr = np.frombuffer(buf, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    nun = C.c_uint64()
    L.check(ctx._lib.umigpu_push_bam_records(ctx._h, len(offsets) - 1, arr.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p),
                                             umi_sep, first_read_index, C.byref(nun)), ctx._h)
    ctx._keepalive.append((arr, offsets))
    return int(nun.value)


def _mate_fields(buf, o: int):
    l_name = buf[o + 12]
    flag, = struct.unpack_from("<H", buf, o + 18)
    tid, pos = struct.unpack_from("<ii", buf, o + 4)
    mtid, mpos = struct.unpack_from("<ii", buf, o + 24)
    return bytes(buf[o + 36: o + 36 + max(l_name - 1, 0)]), flag, tid, pos, mtid, mpos


def select_mates(buf, offsets, kept) -> list:
    """UcWriter::write / write_reversed, deduplicate_sam.rs:382-462: record numbers of the mates of the kept reads.
    Every kept paired read registers (qname, mate ref, mate pos); every mapped, paired, last-in-template record with a
    mapped mate whose (qname, ref, pos) is registered is written and the entry removed (a repeated mate record is written
    once, the first in file order)."""
    want = set()
    for i in kept:
        name, flag, _tid, _pos, mtid, mpos = _mate_fields(buf, int(offsets[i]))
        if flag & 0x1:
            want.add((name, mtid, mpos))
    out = []
    for i in range(len(offsets) - 1):
        name, flag, tid, pos, _mtid, _mpos = _mate_fields(buf, int(offsets[i]))
        if (flag & 0x4) or not (flag & 0x1) or not (flag & 0x80) or (flag & 0x8):
            continue
        key = (name, tid, pos)
        if key in want:
            want.discard(key)
            out.append(i)
    return out


def _passes_filters(buf, o: int, args) -> bool:
    """deduplicate_sam.rs:96-129: does this record reach UcSAMRead::new (:152)?"""
    flag, = struct.unpack_from("<H", buf, o + 18)
    if args is not None and args.paired:
        if (flag & 0x1) and (flag & 0x80):
            return False
        if flag & 0x4:
            return False
        if not (flag & 0x1):
            return not args.remove_unpaired
        if flag & 0x8:
            return False
        tid, = struct.unpack_from("<i", buf, o + 4); mtid, = struct.unpack_from("<i", buf, o + 24)
        return not (tid != mtid and args.remove_chimeric)
    return not (flag & 0x4)


def autodetect_umi_length(buf, offsets, sep: int, args=None) -> int:
    """utils/read.rs:65-75,87-94 on the first mapped record: the caseless regex ^(?:.*?)SEP([ATCGN]+)(?:.*?)$ —
    i.e. the run of [ATCGN] after the FIRST separator that is followed by at least one such letter.  (get_umi
    itself always cuts after the first separator, utils/read.rs:100-101; when the two disagree the reference
    goes on to panic in to_bitset, and so does this path with UMIGPU_ERR_BAD_BASE.)"""
    letters = b"ACGTNacgtn"
    for i in range(len(offsets) - 1):
        o = int(offsets[i])
        if not _passes_filters(buf, o, args):
            continue
        l_name = buf[o + 12]
        name = bytes(buf[o + 36: o + 36 + l_name - 1])
        p = -1
        while True:
            p = name.find(bytes([sep]), p + 1)
            if p < 0:
                raise ValueError("failed to get the umi")          # regex does not match: unwrap() panics
            if p + 1 < len(name) and name[p + 1] in letters:
                break
        n = 0
        while p + 1 + n < len(name) and name[p + 1 + n] in letters:
            n += 1
        return n
    return 0


def deduplicate_and_merge(args: Cli, device: int = 0, chunk_records: int = 1 << 22) -> dict:
    """BAM in -> BAM out, every CLI flag of the reference that reaches the single-end path honoured
    (-k -u -p --umi_sep --algo --merge --keep-unmapped --paired --remove-unpaired --remove-chimeric; --data ignored
    like the reference).  Survivors are written in input order (canonical), unmapped reads too with --keep-unmapped
    (deduplicate_sam.rs:102-108); with --paired the kept reads' mates follow UcWriter::write_reversed (:409-462)."""
    algo, merge = resolve_cli(args)
    buf = bgzf_read_all(args.input)
    header, _names, first = parse_header(buf)
    offsets, consumed = record_offsets(buf, first)
    n = len(offsets) - 1
    umi_len = args.umi_length or autodetect_umi_length(buf, offsets, args.umi_separator, args)
    if n == 0 or umi_len == 0:
        bgzf_write_all(args.output, header)
        return dict(total_reads=n, n_kept=0)
    fl = 0
    if args.paired:
        fl = L.FLAG_PAIRED | (L.FLAG_REMOVE_UNPAIRED if args.remove_unpaired else 0) | (L.FLAG_REMOVE_CHIMERIC if args.remove_chimeric else 0)
    with Context(umi_len, args.k, args.percentage, algo, merge, device, fl) as ctx:
        for s in range(0, n, chunk_records):
            e = min(n, s + chunk_records)
            push_bam(ctx, buf, offsets[s: e + 1], args.umi_separator, s)
        kept, _, ctr = ctx.finish()
    keep = np.zeros(n, bool)
    keep[kept.astype(np.int64)] = True
    if args.keep_unmapped:
        flags = np.array([struct.unpack_from("<H", buf, int(o) + 18)[0] for o in offsets[:-1]], dtype=np.uint16)
        keep |= (flags & 4) != 0
    if args.paired:
        keep[select_mates(buf, offsets, kept.astype(np.int64))] = True
    out = [header]
    mv = memoryview(buf)
    for i in np.nonzero(keep)[0]:
        out.append(mv[int(offsets[i]): int(offsets[i + 1])])
    bgzf_write_all(args.output, b"".join(out))
    return ctr
