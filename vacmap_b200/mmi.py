"""minimap2 `.mmi` index files (format of minimap2 2.29 `index.c`: `mm_idx_dump` / `mm_idx_load`), SURVEY 8f-3.

The reference builds `<ref>.w{w}_k{k}.mmi` with `minimap2 -d` and hands it to `vacmap_index.Aligner` (vacmap:324-344).
Here the index is built on the GPU in a fraction of the time a file read takes, so an existing `.mmi` is used for what
it uniquely holds -- the sequences and their names -- and `write_mmi` stores a freshly built index in the same format, so
the file stays usable by minimap2 / mappy and by the reference itself.

Layout (little endian):
    "MMI\\2" | u32 w, k, b, n_seq, flag | n_seq x (u8 name length, name, u32 length)
    2^b buckets: i32 n, u64 p[n] | u32 size, size x (u64 key, u64 value)
        a minimizer hash x lives in bucket x & (2^b - 1) under key (x >> b) << 1 | single;
        single: value = y of its only occurrence; else value = start_in_p << 32 | count, occurrences p[start ...], ascending
        y = rid << 32 | last-base position << 1 | strand
    u32 S[(sum_len + 7) / 8]: 4 bits per base (0-3 = ACGT, 4 = N), base o in word o >> 3, shifted by (o & 7) * 4
minimap2 is not in this image and the reference ships no `.mmi`, so the format is restated from the published source and
checked here by round trips only (reader(writer(x)) == x); cross-checking against a real file is left to the first
machine that has one.
"""
import struct

import numpy as np

MAGIC = b"MMI\x02"
_NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _NT4[ord(_c)] = _i
    _NT4[ord(_c.lower())] = _i
_NT4[ord("U")] = 3
_NT4[ord("u")] = 3
_LETTERS = np.frombuffer(b"ACGTNNNNNNNNNNNN", dtype=np.uint8)


def is_mmi(path):
    try:
        with open(path, "rb") as f:
            return f.read(4) == MAGIC
    except OSError:
        return False


def read_mmi(path, with_minimizers=False):
    """-> dict(w, k, b, flag, names, lens, seqs [, keys, counts, occ]); seqs are upper-case ACGTN strings.
    with_minimizers: also the distinct hashes (ascending), their counts and their occurrences as GLOBAL
    (position << 1 | strand) values, the layout `vm_index_minimizers` hands out."""
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise ValueError("%s is not a minimap2 index (bad magic)" % path)
        w, k, b, n_seq, flag = struct.unpack("<5I", f.read(20))
        names, lens = [], []
        for _ in range(n_seq):
            ln = f.read(1)[0]
            names.append(f.read(ln).decode())
            lens.append(struct.unpack("<I", f.read(4))[0])
        lens = np.array(lens, dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        hk, hc, ho = [], [], []
        for i in range(1 << b):
            n = struct.unpack("<i", f.read(4))[0]
            p = np.frombuffer(f.read(8 * n), dtype="<u8") if n else np.zeros(0, np.uint64)
            size = struct.unpack("<I", f.read(4))[0]
            kv = np.frombuffer(f.read(16 * size), dtype="<u8").reshape(-1, 2) if size else np.zeros((0, 2), np.uint64)
            if not with_minimizers or size == 0:
                continue
            key, val = kv[:, 0], kv[:, 1]
            single = (key & np.uint64(1)).astype(bool)
            x = ((key >> np.uint64(1)) << np.uint64(b)) | np.uint64(i)
            cnt = np.where(single, 1, (val & np.uint64(0xffffffff)).astype(np.int64)).astype(np.int64)
            start = (val >> np.uint64(32)).astype(np.int64)
            order = np.argsort(x, kind="stable")
            for j in order:
                hk.append(int(x[j]))
                hc.append(int(cnt[j]))
                ho.append(val[j:j + 1] if single[j] else p[start[j]:start[j] + cnt[j]])
        seqs = None
        if not (flag & 2):
            total = int(starts[-1])
            S = np.frombuffer(f.read(4 * ((total + 7) // 8)), dtype=np.uint8)
            codes = np.empty(S.size * 2, dtype=np.uint8)
            codes[0::2] = S & 15
            codes[1::2] = S >> 4
            letters = _LETTERS[codes[:total]]
            seqs = [letters[starts[i]:starts[i + 1]].tobytes().decode() for i in range(n_seq)]
    out = {"w": w, "k": k, "b": b, "flag": flag, "names": names, "lens": lens, "seqs": seqs}
    if with_minimizers:
        order = np.argsort(np.array(hk, dtype=np.uint64), kind="stable") if hk else np.zeros(0, np.int64)
        keys = np.array(hk, dtype=np.uint64)[order]
        counts = np.array(hc, dtype=np.int32)[order]
        occ_y = np.concatenate([ho[j] for j in order]) if hk else np.zeros(0, np.uint64)
        rid = (occ_y >> np.uint64(32)).astype(np.int64)
        low = occ_y & np.uint64(0xffffffff)
        gpos = (low >> np.uint64(1)).astype(np.int64) + starts[rid]
        out.update(keys=keys, counts=counts, occ=(gpos.astype(np.uint64) << np.uint64(1)) | (low & np.uint64(1)))
    return out


def write_mmi(path, names, seqs, w, k, keys, counts, occ, b=14):
    """Store an index in minimap2's format.  seqs: the sequences (any case; non-ACGT becomes N, as minimap2 stores it);
    keys / counts / occ: the distinct minimizer hashes (ascending), their occurrence counts and the occurrences as GLOBAL
    (last-base position << 1 | strand), key after key (`Index.minimizers()`)."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    keys = np.asarray(keys, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.int64)
    occ = np.asarray(occ, dtype=np.uint64)
    kstart = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    # occurrences as minimap2's y = rid << 32 | local position << 1 | strand
    gpos = (occ >> np.uint64(1)).astype(np.int64)
    rid = np.searchsorted(starts, gpos, side="right") - 1
    y = (rid.astype(np.uint64) << np.uint64(32)) | ((gpos - starts[rid]).astype(np.uint64) << np.uint64(1)) | (occ & np.uint64(1))
    mask = np.uint64((1 << b) - 1)
    bucket = (keys & mask).astype(np.int64)
    order = np.argsort(bucket, kind="stable")          # keys stay ascending inside a bucket
    bstart = np.searchsorted(bucket[order], np.arange((1 << b) + 1))
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<5I", w, k, b, len(names), 0))
        for n, ln in zip(names, lens):
            nb = n.encode()[:255]
            f.write(bytes([len(nb)]) + nb + struct.pack("<I", int(ln)))
        for i in range(1 << b):
            idx = order[bstart[i]:bstart[i + 1]]
            if idx.size == 0:
                f.write(struct.pack("<i", 0) + struct.pack("<I", 0))
                continue
            cnt = counts[idx]
            single = cnt == 1
            multi_cnt = np.where(single, 0, cnt)
            pstart = np.concatenate([[0], np.cumsum(multi_cnt)[:-1]]).astype(np.int64)
            n_p = int(multi_cnt.sum())
            if n_p:
                mi = idx[~single]
                take = np.concatenate([np.arange(kstart[j], kstart[j + 1]) for j in mi])
                p = y[take]
            else:
                p = np.zeros(0, np.uint64)
            key = ((keys[idx] >> np.uint64(b)) << np.uint64(1)) | single.astype(np.uint64)
            val = np.where(single, y[kstart[idx]], (pstart.astype(np.uint64) << np.uint64(32)) | cnt.astype(np.uint64))
            f.write(struct.pack("<i", n_p))
            f.write(p.astype("<u8").tobytes())
            f.write(struct.pack("<I", int(idx.size)))
            f.write(np.stack([key, val], axis=1).astype("<u8").tobytes())
        total = int(starts[-1])
        codes = np.full(((total + 7) // 8) * 8, 0, dtype=np.uint8)
        o = 0
        for s in seqs:
            raw = np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8)
            codes[o:o + raw.size] = _NT4[raw]
            o += raw.size
        f.write((codes[0::2] | (codes[1::2] << 4)).astype(np.uint8).tobytes())
