"""ctypes binding of libvacmap_b200.so (see include/vacmap_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvacmap_b200.so")

VM_OK = 0
NOPRE = -9999999
ABI_VERSION = 9

_ERRS = {-1: "CUDA error", -2: "no CUDA device (vacmap_b200 has no CPU fallback)", -3: "bad argument",
         -4: "out of memory", -5: "bad call order"}


class VacmapB200Error(RuntimeError):
    pass


class ChainParamsC(ctypes.Structure):
    _fields_ = [("kmersize", ctypes.c_int32), ("skipcost", ctypes.c_double), ("maxdiff", ctypes.c_int32),
                ("maxgap", ctypes.c_int32), ("max_factor", ctypes.c_int32), ("fast_t", ctypes.c_int32),
                ("large_readgap", ctypes.c_int32), ("variant", ctypes.c_int32)]


_lib = None


def load():
    """Load the CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VacmapB200Error(
            "libvacmap_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python vacmap_b200/build.py`; there is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.vm_abi_version.restype = i32
    if L.vm_abi_version() != ABI_VERSION:
        raise VacmapB200Error("libvacmap_b200.so ABI %d != binding %d: rebuild" % (L.vm_abi_version(), ABI_VERSION))
    L.vm_ctx_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.vm_ctx_destroy.argtypes = [vp]
    L.vm_ctx_destroy.restype = None
    L.vm_last_error.argtypes = [vp]
    L.vm_last_error.restype = ctypes.c_char_p
    L.vm_kernel_launches.argtypes = [vp]
    L.vm_kernel_launches.restype = i64
    L.vm_set_tables.argtypes = [vp, vp, i64, vp, i64, vp, i64]
    L.vm_host_alloc.argtypes = [vp, i64, ctypes.POINTER(vp)]
    L.vm_host_free.argtypes = [vp, vp]
    L.vm_host_register.argtypes = [vp, vp, i64]
    L.vm_host_unregister.argtypes = [vp, vp]
    L.vm_chain_global_batch.argtypes = [vp, ctypes.POINTER(ChainParamsC), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.vm_chain_global_upload.argtypes = [vp, ctypes.POINTER(ChainParamsC), i64, vp, vp, vp]
    L.vm_chain_global_run.argtypes = [vp, vp]
    L.vm_chain_global_download.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.vm_chain_global_times.argtypes = [vp, vp]
    _lib = L
    return L


def check(ctx, rc):
    if rc != VM_OK:
        msg = ""
        if ctx:
            msg = (load().vm_last_error(ctx) or b"").decode()
        raise VacmapB200Error("%s: %s" % (_ERRS.get(rc, "error %d" % rc), msg))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Context:
    """One per GPU / host thread (vm_ctx)."""

    def __init__(self, device=0):
        L = load()
        h = ctypes.c_void_p()
        rc = L.vm_ctx_create(int(device), ctypes.byref(h))
        if rc != VM_OK:
            raise VacmapB200Error("vm_ctx_create(device=%d): %s" % (device, _ERRS.get(rc, rc)))
        self.h = h
        self.device = device
        from .tables import score_tables
        t = score_tables()
        check(self.h, L.vm_set_tables(self.h, ptr(t.extra), len(t.extra), ptr(t.readgapcost), len(t.readgapcost),
                                      ptr(t.log2cache), len(t.log2cache)))

    @property
    def kernel_launches(self):
        return int(load().vm_kernel_launches(self.h))

    def pinned_bytes(self, n):
        """uint8 array of n bytes in page-locked host memory (vm_host_alloc): a read batch packed into it is copied to
        the device by DMA, overlapped with the other sub-batches' kernels.  Freed with the array."""
        import weakref
        import numpy as np
        L = load()
        p = ctypes.c_void_p()
        check(self.h, L.vm_host_alloc(self.h, int(n), ctypes.byref(p)))
        buf = (ctypes.c_uint8 * max(int(n), 1)).from_address(p.value)
        weakref.finalize(buf, L.vm_host_free, None, p)
        return np.frombuffer(buf, dtype=np.uint8)[:int(n)]

    def close(self):
        if getattr(self, "h", None):
            load().vm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
