"""vacmap_b200 -- B200-native implementation of VACmap's per-read alignment hot path.

Host side is Python (as in the reference) over a thin ctypes binding of the C ABI in
``include/vacmap_b200.h`` (``libvacmap_b200.so``: hand-written CUDA for sm_100a).
There is no CPU fallback: every operator raises if the library or a CUDA device is
missing.
"""
from . import _lib  # noqa: F401
from .tables import score_tables  # noqa: F401
from .chain import ChainParams, chain_global_batch, chain_local_batch, GlobalChainer  # noqa: F401
from .align import Index, Aligner, Record, default_option, read_fastx  # noqa: F401

__all__ = ["score_tables", "ChainParams", "chain_global_batch", "chain_local_batch", "GlobalChainer", "Index", "Aligner", "Record",
           "default_option", "read_fastx"]
