"""BAM in and out without samtools / pysam (neither is in this image).

The reference reads unaligned BAM through `pysam.AlignmentFile(path, check_sq=False)` (`vacmap:439-466`: name, sequence,
qualities; a record with FLAG 16 is reverse-complemented back to the read's own strand) and writes BAM by piping its SAM
text into `samtools view -b` / `samtools sort` (`output_functions.py:200-222`).  Here the same SAM lines the emitter
produces (`sam.py`) are encoded as BAM records and written as BGZF blocks directly (SAM spec v1, sections 4.1-4.2), and
a BAM reader yields what the reference takes from pysam.  `*.sorted.bam` is sorted by (reference, position) in memory like
`samtools sort` orders it; the `--write-index` side file (.csi) is not produced.
"""
import gzip
import struct
import zlib

import numpy as np

_SEQ_CODES = "=ACMGRSVTWYHKDBN"
_SEQ_ENC = {c: i for i, c in enumerate(_SEQ_CODES)}
_CIGAR_OPS = "MIDNSHP=X"
_CIGAR_ENC = {c: i for i, c in enumerate(_CIGAR_OPS)}
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_COMP = str.maketrans("ACGTUNRYKMBDHVacgtunrykmbdhv", "TGCAANYRMKVHDBtgcaanyrmkvhdb")


# per-base work goes through numpy tables: nibble codes of the letters, and the two letters of every packed byte
_ENC_TAB = np.full(256, 15, dtype=np.uint8)
for _c, _i in _SEQ_ENC.items():
    _ENC_TAB[ord(_c)] = _i
    _ENC_TAB[ord(_c.lower())] = _i
_DEC_TAB = np.array([[ord(_SEQ_CODES[b >> 4]), ord(_SEQ_CODES[b & 15])] for b in range(256)], dtype=np.uint8)


def _unpack_seq(packed, l_seq):
    return _DEC_TAB[np.frombuffer(packed, dtype=np.uint8)].reshape(-1)[:l_seq].tobytes().decode()


def _pack_seq(seq):
    codes = _ENC_TAB[np.frombuffer(seq.encode(), dtype=np.uint8)]
    if len(codes) & 1:
        codes = np.append(codes, np.uint8(0))
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()


# ---------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------
def read_bam(path):
    """-> (name, sequence, quality string or None) per record, in file order, as the reference takes them from pysam
    (`vacmap:455-470`): sequence upper-case on the read's own strand (FLAG 16 records are reverse-complemented, their
    qualities reversed); records without a sequence are skipped."""
    with gzip.open(path, "rb") as f:          # BGZF is a series of gzip members
        def need(n):
            b = f.read(n)
            if len(b) != n:
                raise EOFError("truncated BAM file: %s" % path)
            return b
        if need(4) != b"BAM\x01":
            raise ValueError("not a BAM file: %s" % path)
        (l_text,) = struct.unpack("<i", need(4))
        need(l_text)
        (n_ref,) = struct.unpack("<i", need(4))
        for _ in range(n_ref):
            (l_name,) = struct.unpack("<i", need(4))
            need(l_name + 4)
        while True:
            head = f.read(4)
            if not head:
                return
            if len(head) != 4:
                raise EOFError("truncated BAM record: %s" % path)
            (block_size,) = struct.unpack("<i", head)
            rec = need(block_size)
            _ref, _pos, l_read_name, _mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", rec, 0)
            p = 32
            name = rec[p:p + l_read_name - 1].decode()
            p += l_read_name + 4 * n_cigar
            if l_seq <= 0:
                continue                      # pysam: query_sequence is None -> the reference skips the read
            packed = rec[p:p + (l_seq + 1) // 2]
            p += (l_seq + 1) // 2
            seq = _unpack_seq(packed, l_seq).upper()
            q = rec[p:p + l_seq]
            qual = None if (not q or q[0] == 0xff) else q
            if flag & 16:
                seq = seq.translate(_COMP)[::-1]
                if qual is not None:
                    qual = qual[::-1]
            yield (name, seq, None if qual is None else (np.frombuffer(qual, dtype=np.uint8) + np.uint8(33)).tobytes().decode("ascii"))


# ---------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------
def reg2bin(beg, end):
    """SAM spec 5.3: the UCSC bin of a zero-based half-open interval."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _parse_cigar(text):
    if text == "*":
        return []
    ops, n = [], 0
    for ch in text:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            ops.append((n, _CIGAR_ENC[ch]))
            n = 0
    return ops


def _aux(tag, typ, val):
    t = tag.encode()
    if typ == "A":
        return t + b"A" + val.encode()
    if typ == "i":
        v = int(val)
        for code, fmt, lo, hi in (("C", "<B", 0, 255), ("c", "<b", -128, 127), ("S", "<H", 0, 65535), ("s", "<h", -32768, 32767),
                                  ("I", "<I", 0, 4294967295), ("i", "<i", -2147483648, 2147483647)):
            if lo <= v <= hi:
                return t + code.encode() + struct.pack(fmt, v)
        raise ValueError("integer tag out of range: %s" % val)
    if typ == "f":
        return t + b"f" + struct.pack("<f", float(val))
    if typ == "Z":
        return t + b"Z" + val.encode() + b"\x00"
    if typ == "H":
        return t + b"H" + val.encode() + b"\x00"
    if typ == "B":
        sub, *items = val.split(",")
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
        conv = float if sub == "f" else int
        return t + b"B" + sub.encode() + struct.pack("<i", len(items)) + struct.pack("<%d%s" % (len(items), fmt), *[conv(x) for x in items])
    raise ValueError("unknown tag type %s" % typ)


def encode_record(line, ref_index):
    """One SAM line -> one BAM alignment record (with its block_size prefix), and its sort key."""
    f = line.rstrip("\n").split("\t")
    qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
    flag, pos, mapq, pnext, tlen = int(flag), int(pos) - 1, int(mapq), int(pnext) - 1, int(tlen)
    ref_id = -1 if rname == "*" else ref_index[rname]
    next_id = -1 if rnext == "*" else (ref_id if rnext == "=" else ref_index[rnext])
    ops = _parse_cigar(cigar)
    l_seq = 0 if seq == "*" else len(seq)
    ref_len = sum(n for n, op in ops if op in (0, 2, 3, 7, 8))
    end = pos + (ref_len if ref_len > 0 else 1)
    cigar_in_tag = len(ops) > 65535            # SAM spec 4.2.2: long CIGARs move to the CG tag
    body = bytearray()
    name = qname.encode() + b"\x00"
    n_cigar = 2 if cigar_in_tag else len(ops)
    body += struct.pack("<iiBBHHHiiii", ref_id, pos, len(name), mapq, reg2bin(max(pos, 0), max(end, 1)), n_cigar, flag, l_seq, next_id, pnext, tlen)
    body += name
    if cigar_in_tag:
        body += struct.pack("<II", l_seq << 4 | 4, ref_len << 4 | 3)
    else:
        body += struct.pack("<%dI" % len(ops), *[n << 4 | op for n, op in ops])
    if l_seq:
        body += _pack_seq(seq)
        body += bytes([0xff]) * l_seq if qual == "*" else (np.frombuffer(qual.encode(), dtype=np.uint8) - np.uint8(33)).tobytes()
    for tag in f[11:]:
        name2, typ, val = tag.split(":", 2)
        body += _aux(name2, typ, val)
    if cigar_in_tag:
        body += b"CGBI" + struct.pack("<i", len(ops)) + struct.pack("<%dI" % len(ops), *[n << 4 | op for n, op in ops])
    key = (ref_id if ref_id >= 0 else 1 << 31, pos)
    return struct.pack("<i", len(body)) + bytes(body), key


class BamWriter:
    """`out.bam`: records in the order written (`samtools view -b`); `out.sorted.bam`: buffered and sorted by
    (reference, position), unmapped last, stable (`samtools sort`; header gets SO:coordinate)."""

    def __init__(self, path, header_text, level=6):
        self.f = open(path, "wb")
        self.level = level
        self.sorted = path.endswith("sorted.bam")
        self.buf = bytearray()
        self.pending = []
        self._pool = None
        refs = []
        lines = header_text.splitlines()
        for ln in lines:
            if ln.startswith("@SQ"):
                d = dict(x.split(":", 1) for x in ln.split("\t")[1:])
                refs.append((d["SN"], int(d["LN"])))
        if self.sorted:
            lines = [("@HD\tVN:1.0\tSO:coordinate" if ln.startswith("@HD") else ln) for ln in lines]
        text = ("\n".join(lines) + "\n").encode()
        self.ref_index = {n: i for i, (n, _) in enumerate(refs)}
        head = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
        for n, ln in refs:
            nm = n.encode() + b"\x00"
            head += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
        self._put(bytes(head))

    @staticmethod
    def _compress(args):
        data, level = args
        comp = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = comp.compress(data) + comp.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cdata) + 25) + cdata +
                struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))

    def _flush_blocks(self, blocks):
        """BGZF blocks are independent deflate streams: compressed by a few threads (zlib releases the GIL, like
        `samtools view -@ 8`), written in order."""
        if not blocks:
            return
        if len(blocks) < 4:
            for b in blocks:
                self.f.write(self._compress((b, self.level)))
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            import os
            self._pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
        for c in self._pool.map(self._compress, [(b, self.level) for b in blocks]):
            self.f.write(c)

    def _put(self, data):
        self.buf += data
        if len(self.buf) >= 64 * 0xff00:           # a few megabytes at a time
            n = len(self.buf) // 0xff00
            view = bytes(self.buf[:n * 0xff00])
            del self.buf[:n * 0xff00]
            self._flush_blocks([view[i * 0xff00:(i + 1) * 0xff00] for i in range(n)])

    def write_sam_lines(self, lines):
        for ln in lines:
            if not ln or ln.startswith("@"):
                continue
            rec, key = encode_record(ln, self.ref_index)
            if self.sorted:
                self.pending.append((key, len(self.pending), rec))
            else:
                self._put(rec)

    def close(self):
        if self.f is None:
            return
        if self.sorted:
            self.pending.sort(key=lambda t: (t[0], t[1]))
            for _, _, rec in self.pending:
                self._put(rec)
            self.pending = []
        if self.buf:
            view = bytes(self.buf)
            self._flush_blocks([view[i:i + 0xff00] for i in range(0, len(view), 0xff00)])
            self.buf = bytearray()
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        self.f.write(_BGZF_EOF)
        self.f.close()
        self.f = None


def read_bam_records(path):
    """Test helper: (header text, [(name, length)], [decoded record dicts]) of a BAM file, every field."""
    out = []
    with gzip.open(path, "rb") as f:
        data = f.read()
    assert data[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].decode()
    p = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, p)
    p += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p)
        nm = data[p + 4:p + 4 + l_name - 1].decode()
        (ln,) = struct.unpack_from("<i", data, p + 4 + l_name)
        refs.append((nm, ln))
        p += 8 + l_name
    while p < len(data):
        (bs,) = struct.unpack_from("<i", data, p)
        rec = data[p + 4:p + 4 + bs]
        p += 4 + bs
        ref_id, pos, l_rn, mapq, bin_, n_cig, flag, l_seq, next_id, pnext, tlen = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        q = 32
        name = rec[q:q + l_rn - 1].decode()
        q += l_rn
        cig = struct.unpack_from("<%dI" % n_cig, rec, q)
        q += 4 * n_cig
        packed = rec[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = _unpack_seq(packed, l_seq)
        qual = rec[q:q + l_seq]
        q += l_seq
        out.append(dict(name=name, flag=flag, ref_id=ref_id, pos=pos, mapq=mapq, bin=bin_, next_id=next_id, pnext=pnext, tlen=tlen,
                        cigar="".join("%d%s" % (c >> 4, _CIGAR_OPS[c & 15]) for c in cig) or "*", seq=seq or "*",
                        qual="*" if (not qual or qual[0] == 0xff) else bytes(b + 33 for b in qual).decode(), aux=bytes(rec[q:])))
    return text, refs, out
