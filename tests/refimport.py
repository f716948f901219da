"""Import the reference's mode modules in THIS container for golden-vector generation.

Test infrastructure only.  /root/reference is absent on the GPU box, so nothing
marked `gpu`, `smoke()` or `bench.py` may import this module.  The absent
third-party packages are stubbed (SURVEY.md section 8c "stub recipe"); the stubs
for `vacmap_index`, `edlib`, `Bio.Seq.Seq` and `cigar.Cigar` can be replaced by
the oracle's C restatements (see `install_native_shims`).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("VACMAP_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "vacmap"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Seq:
    """Bio.Seq.Seq stand-in: upper-case ACGTN reverse complement only."""
    _T = str.maketrans("ACGTNacgtn", "TGCANtgcan")

    def __init__(self, s):
        self.s = str(s)

    def reverse_complement(self):
        return _Seq(self.s.translate(self._T)[::-1])

    def __str__(self):
        return self.s


class _Cigar:
    """cigar.Cigar stand-in: len() = query bases consumed (M, I, S, =, X)."""

    def __init__(self, s):
        self.s = s

    def __len__(self):
        n = 0
        num = 0
        for ch in self.s:
            if ch.isdigit():
                num = num * 10 + ord(ch) - 48
            else:
                if ch in "MIS=X":
                    n += num
                num = 0
        return n


def install_stubs():
    _stub("edlib")
    _stub("pysam")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    bio = _stub("Bio")
    seqio = _stub("Bio.SeqIO")
    seq = _stub("Bio.Seq", Seq=_Seq)
    bio.SeqIO = seqio
    bio.Seq = seq
    _stub("cigar", Cigar=_Cigar)
    _stub("vacmap_index")
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)


_cache = {}


def load_mode(mode="clrnano"):
    """Return the imported reference module `vacmap.mammap_<mode>`."""
    if mode in _cache:
        return _cache[mode]
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    install_stubs()
    import importlib
    mod = importlib.import_module("vacmap.mammap_" + mode)
    _cache[mode] = mod
    return mod
