"""`.mmi` reader / writer (vacmap_b200/mmi.py, format of minimap2 2.29 index.c): round trips on a synthetic index --
names, sequences (non-ACGT -> N, the 4-bit packing), single- and multi-occurrence minimizers across buckets and contigs.
There is no `.mmi` fixture to pin the format against (minimap2 is not in the image, the reference ships none)."""
import numpy as np

from vacmap_b200 import mmi


def test_mmi_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    names = ["chr1", "chrUn_x y", "c3"]
    seqs = ["".join(rng.choice(list("ACGTacgtNRY"), size=n)) for n in (1003, 17, 260)]
    lens = np.array([len(s) for s in seqs])
    starts = np.concatenate([[0], np.cumsum(lens)])
    k = 15
    keys = np.unique(rng.integers(0, 1 << 30, size=400).astype(np.uint64))
    counts = rng.choice([1, 1, 1, 2, 3, 7], size=len(keys)).astype(np.int32)
    occ = []
    for c in counts:
        gp = np.sort(rng.choice(int(starts[-1]), size=c, replace=False))
        occ.append((gp.astype(np.uint64) << np.uint64(1)) | rng.integers(0, 2, size=c).astype(np.uint64))
    occ = np.concatenate(occ)
    p = str(tmp_path / "x.w10_k15.mmi")
    mmi.write_mmi(p, names, seqs, 10, k, keys, counts, occ)
    assert mmi.is_mmi(p)
    m = mmi.read_mmi(p, with_minimizers=True)
    assert (m["w"], m["k"], m["b"], m["flag"]) == (10, 15, 14, 0)
    assert m["names"] == names and list(m["lens"]) == list(lens)
    norm = [s.upper().translate(str.maketrans("RYN", "NNN")) for s in seqs]
    assert m["seqs"] == norm
    assert (m["keys"] == keys).all() and (m["counts"] == counts).all() and (m["occ"] == occ).all()
    # header bytes as minimap2 lays them out
    raw = open(p, "rb").read()
    assert raw[:4] == b"MMI\x02" and np.frombuffer(raw[4:24], dtype="<u4").tolist() == [10, 15, 14, 3, 0]
    assert raw[24] == 4 and raw[25:29] == b"chr1" and int.from_bytes(raw[29:33], "little") == 1003
