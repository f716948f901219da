"""Host mirror of the reference's per-read alignment interface, over the CUDA library.

``Index`` stands where the reference uses ``vacmap_index.Aligner(path, w=, k=)`` (``vacmap:344``);
``Aligner.align_batch`` stands where the worker loop calls ``get_readmap_DP_test`` for every read
(``mammap_clrnano.py:24117``) and returns, per read, the same ``onemapinfolist`` rows
``(readid, contig, strand, q_st, q_en, r_st, r_en, mapq, cigar)`` (``:20760``).
"""
import collections
import ctypes
import gzip
import weakref

import numpy as np

from . import _lib

OPS = "MIDNSHP=X"

Record = collections.namedtuple("Record", "readid contig strand q_st q_en r_st r_en mapq cigar")


class AlignParamsC(ctypes.Structure):
    _fields_ = [("global_skipcost", ctypes.c_double), ("local_skipcost", ctypes.c_double),
                ("maxdivergence", ctypes.c_double), ("accept_score", ctypes.c_double)] + \
               [(n, ctypes.c_int32) for n in ("global_maxdiff", "local_maxdiff", "check_num", "eqx", "hardclip", "nodiscard",
                                              "max_guides", "local_maxgap", "clamp40", "host_threads", "workers",
                                              "chunk_reads")]


class RecordC(ctypes.Structure):
    _fields_ = [("contig", ctypes.c_int32), ("strand", ctypes.c_int32), ("q_st", ctypes.c_int64), ("q_en", ctypes.c_int64),
                ("r_st", ctypes.c_int64), ("r_en", ctypes.c_int64), ("mapq", ctypes.c_int32), ("cigar_len", ctypes.c_int32),
                ("cigar_off", ctypes.c_int64)]


RECORD_DTYPE = np.dtype([("contig", "<i4"), ("strand", "<i4"), ("q_st", "<i8"), ("q_en", "<i8"), ("r_st", "<i8"),
                         ("r_en", "<i8"), ("mapq", "<i4"), ("cigar_len", "<i4"), ("cigar_off", "<i8")])

# per-mode constants that differ between mammap_clrnano (H), mammap_ccs (L), mammap_sensitive (S)
MODE_CONST = {"H": dict(accept=60.0, max_guides=5, local_maxgap=99, clamp40=0),
              "L": dict(accept=40.0, max_guides=3, local_maxgap=50, clamp40=1),
              "S": dict(accept=40.0, max_guides=0, local_maxgap=99, clamp40=0)}


def default_option(mode="H", **over):
    """The `pdict` the reference CLI builds (vacmap:177-296) -- note the load-bearing `golbal_` spelling."""
    skips = {"L": (59., 40., 0.1), "H": (40., 40., 0.2)}.get(mode, (30., 30., 0.5))
    opt = {"mode": mode, "c": 100, "eqx": False, "md": False, "cigar2cg": False, "copycomments": False, "H": False,
           "fakecigar": False, "Q": False, "debug": False, "shortcs": True, "rg-id": "1", "local_kmersize": 9,
           "local_skipcost": skips[0], "golbal_skipcost": skips[1], "maxdivergence": skips[2],
           "golbal_maxdiff": 50, "local_maxdiff": 30, "markunbalancetra": mode in ("L", "H"),
           "nodiscard": mode not in ("L", "H")}
    opt.update(over)
    return opt


def read_fastx(path, read_comment=False):
    """FASTA / FASTQ(.gz) reader with the `vacmap_index.fastx_read` tuple contract (vacmap:445; kseq behind it):
    multi-line records, a FASTQ record's quality runs until it is as long as the sequence -- so a record with an
    EMPTY sequence (common after trimming) is a zero-length read followed by an intact next record."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rt") as f:
        name, comment, seq, qual, mode = None, None, [], [], None
        nseq = nqual = 0
        for line in f:
            line = line.rstrip("\r\n")
            if mode == "qual":
                # inside a quality string nothing is a header: '@' and '>' are quality characters
                qual.append(line)
                nqual += len(line)
                if nqual >= nseq:
                    mode = "fq_done"
                continue
            if not line:
                continue
            if line[0] in ">@":
                if name is not None:
                    yield _rec(name, comment, seq, qual, read_comment)
                hdr = line[1:].split(None, 1)
                name, comment = (hdr[0] if hdr else ""), (hdr[1] if len(hdr) > 1 else None)
                seq, qual, mode = [], [], ("fa" if line[0] == ">" else "fq")
                nseq = nqual = 0
            elif line[0] == "+" and mode == "fq":
                # an empty sequence has an empty quality: the record is complete at once
                mode = "qual" if nseq > 0 else "fq_done"
            elif mode in ("fa", "fq"):
                seq.append(line)
                nseq += len(line)
        if name is not None:
            yield _rec(name, comment, seq, qual, read_comment)


def _rec(name, comment, seq, qual, read_comment):
    s = "".join(seq)
    q = "".join(qual) if qual else None
    return (name, s, q, comment) if read_comment else (name, s, q)


def _declare(L):
    if getattr(L, "_align_declared", False):
        return
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_index_create.argtypes = [vp, i32, vp, vp, vp, i32, i32, ctypes.POINTER(vp)]
    L.vm_index_destroy.argtypes = [vp]
    L.vm_index_destroy.restype = None
    L.vm_index_arrays.argtypes = [vp, vp, vp, vp]
    L.vm_index_adopt.argtypes = [vp, i32, vp, vp, vp, vp, vp, ctypes.POINTER(vp)]
    L.vm_index_minimizers.argtypes = [vp, vp, vp, vp]
    L.vm_index_info.argtypes = [vp] + [vp] * 6
    L.vm_index_contig.argtypes = [vp, i32, vp, vp, vp, vp]
    L.vm_align_batch.argtypes = [vp, vp, ctypes.POINTER(AlignParamsC), i64, vp, vp, ctypes.POINTER(vp)]
    L.vm_align_resident.argtypes = [vp, vp, ctypes.POINTER(AlignParamsC), i64, vp, vp, ctypes.POINTER(vp)]
    L.vm_align_submit.argtypes = [vp, vp, ctypes.POINTER(AlignParamsC), i64, vp, vp, i32, ctypes.POINTER(vp)]
    L.vm_align_wait.argtypes = [vp, ctypes.POINTER(vp)]
    L.vm_reads_upload.argtypes = [vp, vp, i64, vp, vp]
    for f in ("vm_result_num_records", "vm_result_num_cigar_ops"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = i64
    for f in ("vm_result_read_offsets", "vm_result_records", "vm_result_cigar"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = vp
    L.vm_result_stage_times.argtypes = [vp]
    L.vm_result_stage_times.restype = ctypes.c_char_p
    L.vm_result_read_status.argtypes = [vp]
    L.vm_result_read_status.restype = vp
    L.vm_result_free.argtypes = [vp]
    L.vm_result_free.restype = None
    L._align_declared = True


class Index:
    """Reference index resident on one GPU: `.k`, `.w`, `.seq_offset`, `.seq(name)` as the reference uses them."""

    def __init__(self, ref, w=10, k=15, ctx=None, device=0):
        """ref: path to a FASTA(.gz) or to a minimap2 `.mmi` (its sequences are used, the index itself is rebuilt on the
        GPU, which is faster than reading it), or a list of (name, sequence)."""
        L = _lib.load()
        _declare(L)
        self.ctx = ctx or _lib.default_context(device)
        if isinstance(ref, (str, bytes)):
            from . import mmi
            if mmi.is_mmi(ref):
                m = mmi.read_mmi(ref)
                if m["seqs"] is None:
                    raise ValueError("%s was written without sequences (minimap2 --idx-no-seq)" % ref)
                contigs = list(zip(m["names"], m["seqs"]))
            else:
                contigs = [(n, s) for n, s, _ in read_fastx(ref)]
        else:
            contigs = list(ref)
        self.k, self.w = int(k), int(w)
        self.names = [n for n, _ in contigs]
        enc = [s.encode() if isinstance(s, str) else bytes(s) for _, s in contigs]
        n = len(contigs)
        names_c = (ctypes.c_char_p * n)(*[x.encode() for x in self.names])
        seqs_c = (ctypes.c_char_p * n)(*enc)
        lens = np.array([len(e) for e in enc], dtype=np.int64)
        h = ctypes.c_void_p()
        _lib.check(self.ctx.h, L.vm_index_create(self.ctx.h, n, names_c, seqs_c, _lib.ptr(lens), self.w, self.k,
                                                 ctypes.byref(h)))
        self._finish(h, lens)

    def _finish(self, h, lens):
        L = _lib.load()
        self.h = h
        self.lens = np.asarray(lens, dtype=np.int64)
        self.starts = np.concatenate([[0], np.cumsum(self.lens)[:-1]]).astype(np.int64)
        info = [ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()]
        L.vm_index_info(self.h, *[ctypes.byref(x) for x in info])
        self.n_minimizers, self.n_keys, self.mid_occ = info[3].value, info[4].value, info[5].value

    def arrays(self):
        """The built index as device arrays: ([(device pointer, bytes)] x 5, meta int64[8]) -- reference, hash table,
        occurrences, 9-mer positions, 9-mer offsets (`vm_index_arrays`); what `adopt` takes on another rank."""
        L = _lib.load()
        ptrs = (ctypes.c_void_p * 5)()
        nbytes = np.zeros(5, np.int64)
        meta = np.zeros(8, np.int64)
        _lib.check(self.ctx.h, L.vm_index_arrays(self.h, ptrs, _lib.ptr(nbytes), _lib.ptr(meta)))
        return [(int(ptrs[i] or 0), int(nbytes[i])) for i in range(5)], meta

    @classmethod
    def adopt(cls, names, lens, ptrs, nbytes, meta, ctx=None, device=0, keep=None):
        """An index over device arrays the caller owns (`vm_index_adopt`); `keep` = whatever must stay alive with it
        (e.g. the torch tensors an NCCL broadcast filled)."""
        L = _lib.load()
        _declare(L)
        self = cls.__new__(cls)
        self.ctx = ctx or _lib.default_context(device)
        self.names = list(names)
        self.w, self.k = int(meta[5]), int(meta[6])
        self._keep = keep
        n = len(self.names)
        names_c = (ctypes.c_char_p * n)(*[x.encode() for x in self.names])
        lens = np.asarray(lens, dtype=np.int64)
        p = (ctypes.c_void_p * 5)(*[int(x) for x in ptrs])
        nb = np.asarray(nbytes, dtype=np.int64)
        mt = np.asarray(meta, dtype=np.int64)
        h = ctypes.c_void_p()
        _lib.check(self.ctx.h, L.vm_index_adopt(self.ctx.h, n, names_c, _lib.ptr(lens), p, _lib.ptr(nb), _lib.ptr(mt), ctypes.byref(h)))
        self._finish(h, lens)
        return self

    def minimizers(self):
        """(distinct hashes ascending uint64[n_keys], counts int32[n_keys], occurrences uint64[n_minimizers] as global
        last-base position << 1 | strand) -- what a `.mmi` stores."""
        L = _lib.load()
        keys = np.zeros(self.n_keys, np.uint64)
        counts = np.zeros(self.n_keys, np.int32)
        occ = np.zeros(self.n_minimizers, np.uint64)
        _lib.check(self.ctx.h, L.vm_index_minimizers(self.h, _lib.ptr(keys), _lib.ptr(counts), _lib.ptr(occ)))
        return keys, counts, occ

    def write_mmi(self, path):
        """Store the index as a minimap2 `.mmi` (`<ref>.w{w}_k{k}.mmi` is where the reference looks, vacmap:326)."""
        from . import mmi
        keys, counts, occ = self.minimizers()
        mmi.write_mmi(path, self.names, [self.seq(n) for n in self.names], self.w, self.k, keys, counts, occ)

    @property
    def seq_offset(self):
        return [(n.encode(), int(l), int(s)) for n, l, s in zip(self.names, self.lens, self.starts)]

    def seq(self, name, start=0, end=0x7fffffff):
        i = self.names.index(name)
        p, ln = ctypes.c_char_p(), ctypes.c_int64()
        _lib.load().vm_index_contig(self.h, i, None, None, ctypes.byref(ln), ctypes.byref(p))
        return ctypes.string_at(p, ln.value).decode()[start:end]

    def close(self):
        if getattr(self, "h", None):
            _lib.load().vm_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Aligner:
    """Batch form of the reference worker loop: reads in, `onemapinfolist` per read out."""

    def __init__(self, index, option=None, mode="H", host_threads=0, workers=0, chunk_reads=0):
        self.index = index
        self.mode = mode
        self.option = option or default_option(mode)
        mc = MODE_CONST[mode]
        o = self.option
        self.params = AlignParamsC(o["golbal_skipcost"], o["local_skipcost"], o["maxdivergence"], mc["accept"],
                                   o["golbal_maxdiff"], o["local_maxdiff"], o["c"], int(o["eqx"]), int(o["H"]),
                                   int(o["nodiscard"]), mc["max_guides"], mc["local_maxgap"], mc["clamp40"], host_threads, workers, chunk_reads)
        self.last_stage_ms = {}
        self.last_status = np.zeros(0, np.int32)     # per-read VM_READ_* codes of the last collected batch

    def upload_reads(self, seq_cat, seq_off):
        """Put a packed batch in HBM ahead of `align_packed(..., resident=True)` (device-resident timing)."""
        L = _lib.load()
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        ctx = self.index.ctx
        _lib.check(ctx.h, L.vm_reads_upload(ctx.h, self.index.h, len(seq_off) - 1, seq_cat, _lib.ptr(seq_off)))

    def submit_packed(self, seq_cat, seq_off, resident=False):
        """Queue a packed batch (vm_align_submit) and return a handle for `wait`; batches submitted back to back keep
        the device busy across batch boundaries.  seq_cat / seq_off are kept alive by the handle."""
        L = _lib.load()
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        n = len(seq_off) - 1
        job = ctypes.c_void_p()
        ctx = self.index.ctx
        if isinstance(seq_cat, np.ndarray):      # e.g. Context.pinned_bytes: used in place (page-locked -> DMA upload)
            if seq_cat.dtype != np.uint8 or not seq_cat.flags.c_contiguous:
                raise ValueError("seq_cat: contiguous uint8 array or bytes")
            buf, arg = seq_cat, _lib.ptr(seq_cat)
        else:
            buf = arg = seq_cat if isinstance(seq_cat, bytes) else bytes(seq_cat)
        _lib.check(ctx.h, L.vm_align_submit(ctx.h, self.index.h, ctypes.byref(self.params), n, arg, _lib.ptr(seq_off),
                                            1 if resident else 0, ctypes.byref(job)))
        return (job, n, buf, seq_off)

    def wait(self, handle):
        """-> (rec_off int64[n+1], records structured array, cigar uint32 array) of a submitted batch."""
        L = _lib.load()
        job, n, _buf, _off = handle
        res = ctypes.c_void_p()
        ctx = self.index.ctx
        _lib.check(ctx.h, L.vm_align_wait(job, ctypes.byref(res)))
        keep = False
        try:
            nrec, nops = L.vm_result_num_records(res), L.vm_result_num_cigar_ops(res)
            off = np.ctypeslib.as_array(ctypes.cast(L.vm_result_read_offsets(res), ctypes.POINTER(ctypes.c_int64)),
                                        shape=(n + 1,)).copy()
            if nrec:
                raw = np.ctypeslib.as_array(ctypes.cast(L.vm_result_records(res), ctypes.POINTER(ctypes.c_uint8)),
                                            shape=(nrec * RECORD_DTYPE.itemsize,))
                recs = raw.view(RECORD_DTYPE).copy()
            else:
                recs = np.zeros(0, dtype=RECORD_DTYPE)
            self.last_status = np.ctypeslib.as_array(ctypes.cast(L.vm_result_read_status(res), ctypes.POINTER(ctypes.c_int32)),
                                                     shape=(n,)).copy() if n else np.zeros(0, np.int32)
            txt = (L.vm_result_stage_times(res) or b"").decode()
            self.last_stage_ms = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in txt.split(";") if "=" in kv}
            if nops:
                # the CIGAR arena (the bulk of the result) is not copied: the array views the library's memory,
                # which is released when the last view of it is gone
                buf = (ctypes.c_uint32 * nops).from_address(L.vm_result_cigar(res))
                weakref.finalize(buf, L.vm_result_free, res)
                keep = True
                cig = np.frombuffer(buf, dtype=np.uint32)
            else:
                cig = np.zeros(0, dtype=np.uint32)
        finally:
            if not keep:
                L.vm_result_free(res)
        return off, recs, cig

    def align_packed(self, seq_cat, seq_off, resident=False):
        """seq_cat: bytes of all (upper-case) reads; seq_off int64[n+1].
        -> (rec_off int64[n+1], records structured array, cigar uint32 array)."""
        return self.wait(self.submit_packed(seq_cat, seq_off, resident=resident))

    def rows_of(self, rid, recs, cig):
        """`onemapinfolist` rows (clrnano:20760) of one read from its slice of the record array."""
        rows = []
        for r in recs:
            ops = cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]]
            rows.append(Record(rid, self.index.names[r["contig"]], "+" if r["strand"] == 1 else "-", int(r["q_st"]),
                               int(r["q_en"]), int(r["r_st"]), int(r["r_en"]), int(r["mapq"]), cigar_string(ops)))
        return rows

    def align_batch(self, reads):
        """reads: list of (readid, sequence).  -> list (per read) of lists of Record."""
        enc = [s.upper().encode() for _, s in reads]
        off = np.zeros(len(reads) + 1, dtype=np.int64)
        for i, e in enumerate(enc):
            off[i + 1] = off[i] + len(e)
        rec_off, recs, cig = self.align_packed(b"".join(enc), off)
        return [self.rows_of(rid, recs[rec_off[i]:rec_off[i + 1]], cig) for i, (rid, _) in enumerate(reads)]

    def sam_lines(self, reads, quals=None):
        """SAM text of a batch: per read, the lines `get_bam_dict_str` (clrnano:20841) writes for its records."""
        from . import sam
        o = self.option
        contig2seq = {n: self.index.seq(n) for n in self.index.names}
        contig2iloc = {n: i for i, n in enumerate(self.index.names)}
        out = []
        for i, ((rid, seq), rows) in enumerate(zip(reads, self.align_batch(reads))):
            out.append(sam.get_bam_dict_str(rows, seq.upper(), quals[i] if quals else None, contig2iloc, contig2seq, o["md"],
                                            o["shortcs"], o["cigar2cg"], o["markunbalancetra"], o) if rows else [])
        return out


def cigar_string(ops):
    return "".join("%d%s" % (int(o) >> 4, OPS[int(o) & 0xf]) for o in ops)


def pairs_batch(kind, targets, queries, eqx=False, ctx=None, device=0, band=None):
    """Stage-level base-level kernels on raw (target, query) string pairs.
    kind 'distance' -> list of int (with `band` = list of half-widths k: exact when <= k, else some value > k); 'extend' -> list of (q_e, t_e); 'fill' -> list of CIGAR strings."""
    L = _lib.load()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_pairs_batch.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    ctx = ctx or _lib.default_context(device)
    n = len(targets)
    tb, qb = [t.encode() for t in targets], [q.encode() for q in queries]
    t_off = np.zeros(n + 1, np.int64)
    q_off = np.zeros(n + 1, np.int64)
    for i in range(n):
        t_off[i + 1] = t_off[i] + len(tb[i])
        q_off[i + 1] = q_off[i] + len(qb[i])
    out0, out1 = np.zeros(n, np.int64), np.zeros(n, np.int64)
    k = {"distance": 0, "extend": 1, "fill": 2}[kind]
    if k == 0 and band is not None:
        k = 3
        out1[:] = np.asarray(band, np.int64)
    cap = int(t_off[-1] + q_off[-1] + 2 * n + 16)
    cig = np.zeros(cap if k == 2 else 1, np.uint32)
    _lib.check(ctx.h, L.vm_pairs_batch(ctx.h, k, int(eqx), n, b"".join(tb), _lib.ptr(t_off), b"".join(qb), _lib.ptr(q_off),
                                       _lib.ptr(out0), _lib.ptr(out1), _lib.ptr(cig)))
    if k in (0, 3):
        return [int(v) for v in out0]
    if k == 1:
        return [(int(a), int(b)) for a, b in zip(out0, out1)]
    res, co = [], 0
    for i in range(n):
        res.append(cigar_string(cig[co:co + int(out0[i])]))
        co += len(tb[i]) + len(qb[i]) + 2
    return res


def seed_batch(index, reads, check_num=100):
    """Stage-level seeding: per read (int64[n,4] anchors after the cluster filter and strand flip, need_reverse)."""
    L = _lib.load()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_seed_batch_rows.argtypes = [vp, vp, i32, i64, vp, vp, vp, i64, vp, vp]
    enc = [s.upper().encode() for s in reads]
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    for i, e in enumerate(enc):
        off[i + 1] = off[i] + len(e)
    cap = max(4096, 2 * int(off[-1]))
    row_off = np.zeros(len(reads) + 1, dtype=np.int64)
    nrev = np.zeros(len(reads), dtype=np.int32)
    while True:
        rows = np.zeros((cap, 4), dtype=np.int64)
        rc = L.vm_seed_batch_rows(index.ctx.h, index.h, check_num, len(reads), b"".join(enc), _lib.ptr(off), _lib.ptr(rows), cap,
                                  _lib.ptr(row_off), _lib.ptr(nrev))
        if rc == -4:
            cap = int(row_off[-1]) + 16
            continue
        _lib.check(index.ctx.h, rc)
        break
    return [(rows[row_off[i]:row_off[i + 1]].copy(), bool(nrev[i])) for i in range(len(reads))]


def local_reseed_batch(index, reads, jobs):
    """Stage-level local re-seeding (``vm_local_reseed_batch``).  reads: oriented sequences; jobs: list of
    ``(read_index, windows [(lo, hi) global], guides int64[m, >=2] sorted by read position, readstart, readend)``.
    -> per job, int64[n, 4] anchors in the reference's emission order."""
    L = _lib.load()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    L.vm_local_reseed_batch.argtypes = [vp, vp, i64, vp, vp, i64] + [vp] * 9 + [vp, i64, vp]
    enc = [s.upper().encode() for s in reads]
    off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    nj = len(jobs)
    job_read = np.array([j[0] for j in jobs], np.int32)
    rs = np.array([j[3] for j in jobs], np.int32)
    re_ = np.array([j[4] for j in jobs], np.int32)
    win_off = np.concatenate([[0], np.cumsum([len(j[1]) for j in jobs])]).astype(np.int64)
    g_off = np.concatenate([[0], np.cumsum([len(j[2]) for j in jobs])]).astype(np.int64)
    wl = np.array([w[0] for j in jobs for w in j[1]] or [0], np.int64)
    wh = np.array([w[1] for j in jobs for w in j[1]] or [0], np.int64)
    gx = np.concatenate([np.asarray(j[2])[:, 0] for j in jobs] or [np.zeros(0)]).astype(np.int32)
    gy = np.concatenate([np.asarray(j[2])[:, 1] for j in jobs] or [np.zeros(0)]).astype(np.int64)
    # room for 4 anchors per scanned read position (a job scans [readstart, readend), not the whole read -- a contig's
    # batches are 100 kb of a multi-megabase read); rows beyond row_off[-1] are never looked at
    cap = max(4096, 4 * int(sum(max(0, int(j[4]) - int(j[3])) for j in jobs)))
    ctx = index.ctx
    while True:
        rows = np.empty((cap, 4), np.int64)
        row_off = np.zeros(nj + 1, np.int64)
        rc = L.vm_local_reseed_batch(ctx.h, index.h, len(reads), b"".join(enc), _lib.ptr(off), nj, _lib.ptr(job_read), _lib.ptr(rs),
                                     _lib.ptr(re_), _lib.ptr(win_off), _lib.ptr(wl), _lib.ptr(wh), _lib.ptr(g_off), _lib.ptr(gx),
                                     _lib.ptr(gy), _lib.ptr(rows), cap, _lib.ptr(row_off))
        if rc == -4:      # VM_ERR_NOMEM: row_off is filled, retry with enough room
            cap = int(row_off[-1]) + 16
            continue
        _lib.check(ctx.h, rc)
        return [rows[row_off[j]:row_off[j + 1]].copy() for j in range(nj)]
