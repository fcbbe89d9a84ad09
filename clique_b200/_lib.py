"""ctypes binding of libclq.so (include/clq.h).  Fails loudly when the library is missing: there is no fallback."""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libclq.so"

CLQ_OK, READ_TOO_LONG, SCORING_NOT_REPRESENTABLE, TRACEBACK_DIVERGED, CIGAR_POOL_FULL, NO_CANDIDATE = range(6)
E_INVALID, E_CUDA, E_NOMEM, E_LIMIT, E_STATE, E_UNSUPPORTED = -1, -2, -3, -4, -5, -6

BAND_MAXLEN, BAND_READLEN, BAND_K, BAND_K_SHIFT = 0, 1, 2, 8
SEARCH_FIXED, SEARCH_EXHAUSTIVE, SEARCH_QUICK = 0 << 2, 1 << 2, 2 << 2
SCORE_ONLY, CONVEX, EXTRACT_TAGS, RUSTBIO = 1 << 4, 1 << 5, 1 << 6, 1 << 7

# every symbol include/clq.h declares (tests/test_abi.py checks the built library exports each one)
SYMBOLS = ["clq_version", "clq_strerror", "clq_device_count", "clq_affine_from_f64", "clq_host_alloc", "clq_host_free",
           "clq_ctx_create", "clq_ctx_destroy", "clq_ctx_last_error", "clq_refs_set", "clq_kmer_index_set", "clq_submit",
           "clq_wait", "clq_upload", "clq_launch", "clq_download", "clq_sync", "clq_slot_stats", "clq_set_option",
           "clq_tags_download", "clq_rustbio_scoring", "clq_pack2", "clq_upload_packed2", "clq_submit_packed2"]


class ClqError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("libclq error %d: %s" % (code, msg))
        self.code = code


class AffineInt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("scale", "match", "mismatch", "special", "oe_in", "e_in", "oe_fin", "e_fin",
                                          "b0", "b1", "max_neg")]


class ConvexInt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("match", "mismatch", "special", "o1", "e1", "o2", "e2", "max_neg")]


class Limits(C.Structure):
    _fields_ = [("max_reads", C.c_uint32), ("max_read_bytes", C.c_uint64), ("max_read_len", C.c_uint32),
                ("max_refs", C.c_uint32), ("max_ref_bytes", C.c_uint64), ("cigar_pool_ops", C.c_uint64),
                ("n_slots", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("score_scaled", C.c_int32), ("ref_index", C.c_uint32), ("cigar_off", C.c_uint32),
                ("cigar_len", C.c_uint32), ("status", C.c_uint32), ("matches", C.c_uint32), ("mismatches", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("dp_ms", C.c_float), ("launches", C.c_uint32), ("dp_launches", C.c_uint32),
                ("cells", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("variant", C.c_uint32),
                ("sub_batches", C.c_uint32), ("pack_retries", C.c_uint32), ("reserved0", C.c_uint32)]


def library_path():
    return os.path.join(PKG_DIR, _LIB_NAME)


_lib = None


def load_library():
    """Load libclq.so from the package directory.  Raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ClqError(E_STATE, "%s not found: build it with `make -C clique_b200/csrc` (nvcc, sm_100a); "
                                "there is no CPU fallback" % path)
    L = C.CDLL(path)
    L.clq_version.restype = C.c_int32
    L.clq_strerror.restype = C.c_char_p
    L.clq_strerror.argtypes = [C.c_int32]
    L.clq_device_count.restype = C.c_int32
    L.clq_affine_from_f64.restype = C.c_int32
    L.clq_affine_from_f64.argtypes = [C.c_double] * 6 + [C.POINTER(AffineInt)]
    L.clq_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.clq_host_free.argtypes = [C.c_void_p]
    L.clq_ctx_create.argtypes = [C.c_int32, C.POINTER(Limits), C.POINTER(C.c_void_p)]
    L.clq_ctx_destroy.argtypes = [C.c_void_p]
    L.clq_ctx_destroy.restype = None
    L.clq_ctx_last_error.restype = C.c_char_p
    L.clq_ctx_last_error.argtypes = [C.c_void_p]
    L.clq_refs_set.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.clq_kmer_index_set.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.clq_submit.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_uint32, C.c_double]
    L.clq_wait.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.clq_upload.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.clq_launch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32, C.c_double]
    L.clq_download.argtypes = [C.c_void_p, C.c_int32]
    L.clq_sync.argtypes = [C.c_void_p, C.c_int32]
    L.clq_slot_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Stats)]
    L.clq_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.clq_rustbio_scoring.argtypes = [C.c_int32] * 4 + [C.POINTER(AffineInt)]
    L.clq_tags_download.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    L.clq_pack2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.clq_upload_packed2.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_uint64, C.c_void_p]
    L.clq_submit_packed2.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
    for name in SYMBOLS:
        f = getattr(L, name)
        if f.restype is C.c_int:
            f.restype = C.c_int32
    _lib = L
    return L
