"""SAM record text for the alignments of one read.

Host mirror of the reference's emitter for the per-read path (``mammap_clrnano.py``):
``get_bam_dict_str`` (:20841-21021), ``get_bam_dict_str_comments`` (:21022-), ``reassign_mapq``
(:11661-11707), ``mergecigar_`` (:4773-4796), ``mergecigar_md_`` / ``get_MD_CSshort`` / ``get_MD_CSlong``
(:19012-19148), ``P_alignmentstring`` (:5391-5424) and ``output_functions.nm_from_cigar`` (:300-349).
Same argument meaning, same text -- including the quirks SURVEY Appendix A lists (tag order = dict
insertion order RG, [CG], SA, NM, MD, cs; MD / cs empty unless the CIGAR uses ``=`` / ``X``; NM under
``--H`` computed at the reference's offsets; ``n_cigar`` counting numbers AND letters).  The per-base Python
loops of the reference (NM over ``M`` runs, reverse complement) are numpy comparisons here.
"""
import ctypes
import re

import numpy as np

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")
# Bio.Seq's ambiguous DNA complement table (IUPAC codes, both cases; U -> A)
_COMP = bytes.maketrans(b"ACGTUNRYKMBVDHSWacgtunrykmbvdhsw", b"TGCAANYRMKVBHDSWtgcaanyrmkvbhdsw")


def reverse_complement(seq):
    """``str(Bio.Seq.Seq(seq).reverse_complement())``, IUPAC ambiguity codes included."""
    return seq.encode().translate(_COMP)[::-1].decode()


def sort_by_length(x):            # :66
    return x[4] - x[3]


def reassign_mapq(onemapinfolist):
    """:11661-11707 -- MAPQ 0 for records off the "main" reference-ordered path (quirk A11)."""
    g_list = [0]
    n = len(onemapinfolist)
    while g_list[-1] < n - 1:
        iloc = g_list[-1]
        test_iloc = iloc
        b = onemapinfolist[iloc]
        b_contig, b_q_st, b_q_en, b_r_st, b_r_en = b[1], b[3], b[4], b[5], b[6]
        hit = False
        while test_iloc + 1 < n:
            test_iloc += 1
            t = onemapinfolist[test_iloc]
            if t[1] != b_contig:
                continue
            if t[2] == "+":
                refgap = t[5] - b_r_en
            else:
                refgap = b_r_st - t[6]
            if abs(refgap) > 100000:
                continue
            if refgap < 10:
                g_list.append(test_iloc)
                hit = True
                break
        if not hit:
            g_list.append(iloc + 1)
    keep = set(g_list)
    out = []
    for iloc, rec in enumerate(onemapinfolist):
        rec = list(rec)
        if iloc not in keep:
            rec[7] = 0
        out.append(rec)
    return out


def mergecigar_(cigarstring):
    """:4773-4796 -- adjacent runs of the same op merged; returns the flat [number, op, number, op ...] list."""
    oplist = []
    preop = None
    for num, op in _CIGAR_RE.findall(cigarstring):
        if op == preop:
            oplist[-2] = str(int(oplist[-2]) + int(num))
        else:
            oplist.append(str(int(num)))
            oplist.append(op)
            preop = op
    return oplist


def _codes(s):
    return np.frombuffer(s.upper().encode(), dtype=np.uint8)


def nm_from_cigar(cigar_string, query_seq, ref_seq):
    """output_functions.py:300-349 -- mismatches in M + I + D + X; S advances the query, H does not (quirk A8)."""
    q, r = _codes(query_seq), _codes(ref_seq)
    nm = q_pos = r_pos = 0
    for num, op in _CIGAR_RE.findall(cigar_string):
        n = int(num)
        if op == "M":
            a, b = q[q_pos:q_pos + n], r[r_pos:r_pos + n]
            if len(a) != n or len(b) != n:
                raise IndexError("string index out of range")      # the reference's per-base loop raises here
            nm += int(np.count_nonzero(a != b))
            q_pos += n
            r_pos += n
        elif op == "I":
            nm += n
            q_pos += n
        elif op == "D":
            nm += n
            r_pos += n
        elif op == "N":
            r_pos += n
        elif op == "S":
            q_pos += n
        elif op == "=":
            q_pos += n
            r_pos += n
        elif op == "X":
            nm += n
            q_pos += n
            r_pos += n
    return nm


def md_cs(oplist, target, query, shortcs=True):
    """get_MD_CSshort / get_MD_CSlong (:19012-19112): MD and cs from an =/X CIGAR; ('', '') at the first M."""
    md, cs = [], []
    refloc = readloc = 0
    preop = ""
    equal_value = 0
    for i in range(1, len(oplist), 2):
        value = int(oplist[i - 1])
        op = oplist[i]
        if op == "X":
            if equal_value > 0:
                md.append(str(equal_value))
            elif preop == "D":
                md.append("0")
            md.append(target[refloc])
            cs.append("*" + (target[refloc] + query[readloc]).lower())
            for j in range(1, value):
                md.append("0" + target[refloc + j])
                cs.append("*" + (target[refloc + j] + query[readloc + j]).lower())
            refloc += value
            readloc += value
            equal_value = 0
        elif op == "=":
            if shortcs:
                cs.append(":" + oplist[i - 1])
            else:
                cs.append("=" + target[refloc:refloc + value].upper())
            refloc += value
            readloc += value
            equal_value += value
        elif op == "D":
            if equal_value > 0:
                md.append(str(equal_value))
            elif preop == "X":
                md.append("0")
            md.append("^" + target[refloc:refloc + value])
            cs.append("-" + target[refloc:refloc + value].lower())
            refloc += value
            equal_value = 0
        elif op == "I":
            cs.append("+" + query[readloc:readloc + value].lower())
            readloc += value
            continue
        elif op in ("S", "H"):
            continue
        else:
            return "", ""
        preop = op
    if equal_value > 0:
        md.append(str(equal_value))
    return "".join(md), "".join(cs)


_FIXED = {"QNAME": 0, "FLAG": 1, "RNAME": 2, "POS": 3, "MAPQ": 4, "CIGAR": 5, "RNEXT": 6, "PNEXT": 7, "TLEN": 8, "SEQ": 9,
          "QUAL": 10}


def _tag(tag, value):
    code = "i" if type(value) is int else "f" if type(value) is float else "Z"
    return tag + ":" + code + ":" + str(value)


def alignment_string(infodict, comments=None, with_comments=False):
    """P_alignmentstring (:5391-5424) / P_alignmentstring_comments (:20686-20730)."""
    infolist = ["*", "4", "*", "0", "255", "*", "*", "0", "0", "*", "*"]
    tags = {"QNAME", "FLAG", "RNAME", "POS", "MAPQ", "CIGAR", "RNEXT", "PNEXT", "TLEN", "SEQ", "QUAL", "SA", "NM", "MD", "cs"}
    for key, value in infodict.items():
        if key in _FIXED:
            infolist[_FIXED[key]] = value
        else:
            infolist.append(_tag(key, value))
            tags.add(key)
    if with_comments and isinstance(comments, str):
        for onecomment in comments.split("\t"):
            info = onecomment.split(":")
            if len(info) == 3 and len(info[0]) == 2 and info[0] not in tags and info[1] in ("A", "i", "f", "Z", "H", "B"):
                infolist.append(onecomment)
                tags.add(info[0])
    return "\t".join(infolist)


def _fake_cigar(item, qlen, clipsyb):
    top = str(item[3]) + clipsyb if item[3] > 0 else ""
    tail = str(qlen - item[4]) + clipsyb if (qlen - item[4]) > 0 else ""
    diff = item[4] - item[3] - item[6] + item[5]
    if diff > 0:
        body = str(item[6] - item[5]) + "M" + str(diff) + "I"
    elif diff < 0:
        body = str(item[4] - item[3]) + "M" + str(abs(diff)) + "D"
    else:
        body = str(item[4] - item[3]) + "M"
    return top + body + tail


def mergecigar_nm_(cigarstring):
    """asm mode (mammap_asm.py:23126-23155): mergecigar_ plus NM = the X / D / I lengths -- counted only where a run
    STARTS (the length merged into a preceding run of the same op is not added; quirk kept)."""
    oplist, edit = [], 0
    preop, prenum = "0", 0
    for num, op in _CIGAR_RE.findall(cigarstring):
        n = int(num)
        if op == preop:
            prenum += n
            oplist[-2] = str(prenum)
        else:
            prenum = n
            oplist.append(str(n))
            oplist.append(op)
            if op in "XDI":
                edit += n
            preop = op
    return oplist, edit


def get_bam_dict_str(mapinfo, query, qual, contig2iloc, contig2seq, md, shortcs, cigar2cg, markunbalancetra, option,
                     comments=None, with_comments=False, asm=False):
    """:20841-21021 -- SAM lines of one read's ``onemapinfolist`` rows
    ``(readid, contig, strand, q_st, q_en, r_st, r_en, mapq, cigar)``; longest query span first = primary
    (stable sort then reverse: among equal spans the later row wins, quirk A12)."""
    if markunbalancetra:
        mapinfo = reassign_mapq(mapinfo)
    else:
        mapinfo = [list(x) for x in mapinfo]
    hardclip = option["H"]
    rc_query = reverse_complement(query)
    mapinfo.sort(key=sort_by_length)
    mapinfo = mapinfo[::-1]
    fakecigar = option["fakecigar"]
    clipsyb = "H" if hardclip else "S"
    nms, mds, css, n_cigars, fakes = [], [], [], [], []
    for item in mapinfo:
        oriented = query if item[2] == "+" else rc_query
        target = contig2seq[item[1]][item[5]:item[6]]
        if not md:
            oplist, edit = mergecigar_nm_(item[-1]) if asm else (mergecigar_(item[-1]), None)
            item[-1] = "".join(oplist)
            nms.append(edit if asm else nm_from_cigar(item[8], oriented, target))
            mds.append(None)
            css.append(None)
        else:
            tmp_query = oriented[item[3]:item[4]]
            oplist, edit = mergecigar_nm_(item[-1]) if asm else (mergecigar_(item[-1]), None)
            mdstring, csstring = md_cs(oplist, target, tmp_query, shortcs)
            item[-1] = "".join(oplist)
            nms.append(edit if asm else nm_from_cigar(item[-1], tmp_query, target))
            mds.append(mdstring)
            css.append(csstring)
        n_cigars.append(len(oplist))
        fakes.append(_fake_cigar(item, len(query), clipsyb) if fakecigar else None)
    have_qual = qual is not None and len(qual) == len(query)
    rc_qual = qual[::-1] if have_qual else None
    out = []
    # asm mode (:22838-22841, :22873-22887): the second-longest record is primary when the longest has MAPQ 1 and it
    # has not; MAPQ is written as 60 (anything non-zero) or 1
    primary_iloc = 1 if (asm and len(mapinfo) > 1 and mapinfo[0][7] == 1 and mapinfo[1][7] != 1) else 0
    mq_of = (lambda v: 60 if v != 0 else 1) if asm else (lambda v: v)
    for iloc, primary in enumerate(mapinfo):
        d = {}
        if "rg-id" in option:
            d["RG"] = option["rg-id"]
        d["QNAME"] = primary[0]
        d["RNAME"] = primary[1]
        base_value = 0 if iloc == primary_iloc else 2048
        d["FLAG"] = str(base_value if primary[2] == "+" else 16 + base_value)
        d["POS"] = str(primary[5] + 1)
        if n_cigars[iloc] > 65535 and cigar2cg:
            d["CG"] = primary[8]
        else:
            d["CIGAR"] = primary[8]
        if len(mapinfo) > 1:
            sa = []
            for t, item in enumerate(mapinfo):
                if t == iloc:
                    continue
                sa.append("".join((item[1], ",", str(item[5] + 1), ",", item[2], ",", fakes[t] if fakecigar else item[8], ",",
                                   str(mq_of(item[7])), ",", str(nms[t]) + ";")))
            d["SA"] = "".join(sa)
        d["MAPQ"] = str(mq_of(primary[7]))
        seq, q = (query, qual) if primary[2] == "+" else (rc_query, rc_qual)
        if not hardclip:
            d["SEQ"] = seq
            if have_qual:
                d["QUAL"] = q
        else:
            d["SEQ"] = seq[primary[3]:primary[4]]
            if have_qual:
                d["QUAL"] = q[primary[3]:primary[4]]
        d["NM"] = nms[iloc]
        if md:
            d["MD"] = mds[iloc]
            d["cs"] = css[iloc]
        out.append(alignment_string(d, comments, with_comments))
    return out


def get_bam_dict_str_comments(mapinfo, query, qual, comments, contig2iloc, contig2seq, md, shortcs, cigar2cg, markunbalancetra,
                              option):
    """:21022- -- as above, with the FASTQ comment's well-formed SAM tags copied over (``--copycomments``)."""
    return get_bam_dict_str(mapinfo, query, qual, contig2iloc, contig2seq, md, shortcs, cigar2cg, markunbalancetra, option,
                            comments=comments, with_comments=True)


def iterator_get_bam_dict_str(mapinfo, query, qual, contig2iloc, contig2seq, md, shortcs, cigar2cg, markunbalancetra, option):
    """asm mode's emitter (mammap_asm.py:22757-22941): as get_bam_dict_str with NM taken from the CIGAR alone
    (`mergecigar_n_nm` / `mergecigar_md_cs_nm`), the primary-record rule and the 60 / 1 MAPQ of that mode; a generator
    like the reference's."""
    yield from get_bam_dict_str(mapinfo, query, qual, contig2iloc, contig2seq, md, shortcs, cigar2cg, markunbalancetra, option,
                                asm=True)


# field order of pysam's AlignmentHeader.from_dict (libcalignmentfile.pyx VALID_HEADER_ORDER), which the reference's
# writer uses to print its header dict (output_functions.py:66-76); pysam is absent here, so this order is restated
RG_ORDER = ("ID", "CN", "SM", "LB", "PU", "PI", "DT", "DS", "PL", "FO", "KS", "PG", "PM", "BC")
PG_ORDER = ("PN", "ID", "VN", "PP", "DS", "CL")


def header_text(contigs, rg=None, command_line=None, version="1.0.2"):
    """The header the reference hands to pysam (vacmap:353-370): ``@HD VN:1.0``, one ``@SQ`` per contig, the read
    group every record's RG:Z tag points to (default ``{"ID": "1", "SM": "sample"}``, vacmap:214-218) and the
    ``@PG`` line; ``contigs`` = [(name, length)], ``rg`` = dict of @RG fields (None: the default group)."""
    if rg is None:
        rg = {"ID": "1", "SM": "sample"}
    elif not isinstance(rg, dict):
        rg = {"ID": str(rg), "SM": "sample"}
    lines = ["@HD\tVN:1.0"]
    lines += ["@SQ\tSN:%s\tLN:%d" % (n, ln) for n, ln in contigs]
    keys = [k for k in RG_ORDER if k in rg] + [k for k in rg if k not in RG_ORDER]
    lines.append("@RG\t" + "\t".join("%s:%s" % (k, rg[k]) for k in keys))
    pg = {"ID": "VACmap", "PN": "VACmap", "VN": version, "CL": command_line if command_line is not None else ""}
    lines.append("@PG\t" + "\t".join("%s:%s" % (k, pg[k]) for k in PG_ORDER if k in pg))
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------------------------------------------
# The same text from the library's host threads (csrc/vm_sam.cu): what the command line uses for whole batches
# ---------------------------------------------------------------------------------------------------------------
class SamOptionsC(ctypes.Structure):          # == vm_sam_options
    _fields_ = [("md", ctypes.c_int32), ("shortcs", ctypes.c_int32), ("cigar2cg", ctypes.c_int32), ("markunbalancetra", ctypes.c_int32),
                ("hardclip", ctypes.c_int32), ("fakecigar", ctypes.c_int32), ("copycomments", ctypes.c_int32), ("asm_mode", ctypes.c_int32),
                ("rg_id", ctypes.c_char_p)]


def _pack(items):
    """list of bytes -> (concatenation, int64 offsets[n+1])"""
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        off[1:] = np.cumsum([len(x) for x in items])
    return b"".join(items), off


def pack_rows(per_read_rows, contig_names):
    """`onemapinfolist` rows (readid, contig, strand, q_st, q_en, r_st, r_en, mapq, CIGAR string) per read ->
    (rec_off, recs, cig) in the layout of `Aligner.wait` (what `batch_text` takes)."""
    from .align import OPS, RECORD_DTYPE
    enc = {c: i for i, c in enumerate(OPS)}
    index = {n: i for i, n in enumerate(contig_names)}
    rec_off = np.zeros(len(per_read_rows) + 1, np.int64)
    recs, cig = [], []
    for i, rows in enumerate(per_read_rows):
        for r in rows:
            ops = [(int(n) << 4) | enc[o] for n, o in _CIGAR_RE.findall(r[8])]
            recs.append((index[r[1]], 1 if r[2] == "+" else -1, r[3], r[4], r[5], r[6], r[7], len(ops), len(cig)))
            cig += ops
        rec_off[i + 1] = len(recs)
    return rec_off, (np.array(recs, dtype=RECORD_DTYPE) if recs else np.zeros(0, RECORD_DTYPE)), np.array(cig, dtype=np.uint32)


class ContigTable:
    """The contigs as `vm_sam_batch` takes them: from an `Index` (pointers to the library's own copy, no conversion) or
    from [(name, sequence)]."""

    def __init__(self, source):
        from . import _lib
        if not (hasattr(source, "h") and getattr(source, "h")) and hasattr(source, "names") and hasattr(source, "seq"):
            source = [(n, source.seq(n)) for n in source.names]      # an index-like object without a library handle
        if hasattr(source, "h") and hasattr(source, "names"):
            L = _lib.load()
            self.names = list(source.names)
            ptrs, lens = [], []
            for i in range(len(self.names)):
                p, ln = ctypes.c_void_p(), ctypes.c_int64()
                L.vm_index_contig(source.h, i, None, None, ctypes.byref(ln), ctypes.byref(p))
                ptrs.append(p.value)
                lens.append(ln.value)
            self._keep = source
        else:
            self.names = [n for n, _ in source]
            self._keep = [s.encode() if isinstance(s, str) else s for _, s in source]
            ptrs = [ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p).value for b in self._keep]
            lens = [len(b) for b in self._keep]
        n = len(self.names)
        self._name_bytes = [x.encode() for x in self.names]
        self.c_names = (ctypes.c_char_p * n)(*self._name_bytes)
        self.c_seqs = (ctypes.c_void_p * n)(*ptrs)
        self.c_lens = np.array(lens, dtype=np.int64)
        self.n = n


def batch_text(reads, rec_off, recs, cig, contigs, option, md=False, shortcs=True, cigar2cg=False, markunbalancetra=False,
               copycomments=False, use_qual=True, threads=0, sink=None, packed_seqs=None, asm=False):
    """SAM lines of a whole batch (`vm_sam_batch`): `reads` = [(name, SEQUENCE_UPPER[, qual[, comment]])] in batch order,
    `rec_off` / `recs` / `cig` as `Aligner.wait` returns them, `contigs` a `ContigTable`.  -> (bytes of all lines,
    int64 offsets[n+1] per read); with `sink` (a binary file object) the text is written to it straight from the
    library's buffer and `None` stands in for the bytes.  `packed_seqs` = (bytes of all upper-case reads, int64 offsets[n+1])
    when the caller has packed the batch already (the sequences in `reads` are then not looked at).  Byte-identical to `get_bam_dict_str` / `get_bam_dict_str_comments` read by read; a read
    on which they raise contributes nothing, as in the reference's worker.  `asm`: the contig mode's emitter
    (`iterator_get_bam_dict_str`)."""
    from . import _lib
    L = _lib.load()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_sam_batch.argtypes = [ctypes.POINTER(SamOptionsC), i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32,
                               ctypes.POINTER(vp)]
    L.vm_text_data.restype = vp
    L.vm_text_data.argtypes = [vp]
    L.vm_text_size.restype = i64
    L.vm_text_size.argtypes = [vp]
    L.vm_text_offsets.restype = vp
    L.vm_text_offsets.argtypes = [vp]
    L.vm_text_free.argtypes = [vp]
    L.vm_text_free.restype = None
    n = len(reads)
    if packed_seqs is not None:
        seqs, seq_off = packed_seqs[0], np.ascontiguousarray(packed_seqs[1], dtype=np.int64)
    else:
        seqs, seq_off = _pack([r[1].encode() if isinstance(r[1], str) else r[1] for r in reads])
    names, name_off = _pack([r[0].encode() for r in reads])
    quals = qual_off = comments = comment_off = None
    if use_qual and any(len(r) > 2 and r[2] is not None for r in reads):
        quals, qual_off = _pack([(r[2].encode() if len(r) > 2 and r[2] is not None else b"") for r in reads])
    if copycomments and any(len(r) > 3 and isinstance(r[3], str) for r in reads):
        comments, comment_off = _pack([(r[3].encode() if len(r) > 3 and isinstance(r[3], str) else b"") for r in reads])
    rg = option.get("rg-id")
    opt = SamOptionsC(int(bool(md)), int(bool(shortcs)), int(bool(cigar2cg)), int(bool(markunbalancetra)), int(bool(option["H"])),
                      int(bool(option["fakecigar"])), int(bool(copycomments)), int(bool(asm)), None if rg is None else str(rg).encode())
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    recs = np.ascontiguousarray(recs)
    cig = np.ascontiguousarray(cig, dtype=np.uint32)
    out = vp()
    rc = L.vm_sam_batch(ctypes.byref(opt), contigs.n, contigs.c_names, contigs.c_seqs, _lib.ptr(contigs.c_lens), n, _lib.ptr(rec_off),
                        _lib.ptr(recs) if len(recs) else None, _lib.ptr(cig) if len(cig) else None, seqs, _lib.ptr(seq_off), names,
                        _lib.ptr(name_off), quals, _lib.ptr(qual_off) if qual_off is not None else None, comments,
                        _lib.ptr(comment_off) if comment_off is not None else None, int(threads), ctypes.byref(out))
    if rc != 0:
        raise _lib.VacmapB200Error("vm_sam_batch failed (%d)" % rc)
    try:
        size = L.vm_text_size(out)
        if sink is not None:
            data = None
            if size:
                sink.write(memoryview((ctypes.c_char * size).from_address(L.vm_text_data(out))))
        else:
            data = ctypes.string_at(L.vm_text_data(out), size) if size else b""
        off = np.ctypeslib.as_array(ctypes.cast(L.vm_text_offsets(out), ctypes.POINTER(ctypes.c_int64)), shape=(n + 1,)).copy()
    finally:
        L.vm_text_free(out)
    return data, off
