"""ctypes view of the C++ host layer's batch dispatcher (libclq_host.so, include/clique_host.hpp):
clique::ShardedAligner::align_reads_span -- the product's `align_reads` loop (alignment_functions.rs:63-257, loop at :135) fed
from an in-memory span of reads in plain (unpinned) host memory, over one or several GPUs of the box.  The Python side only
passes pointers; batching, staging into page-locked buffers, submit / wait and the result copy-out all run in C++ threads."""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from ._lib import ClqError
from .aligner import _RESULT_DT, BatchResult

HOST_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libclq_host.so")


class SpanOptions(C.Structure):
    _fields_ = [("max_reads", C.c_uint32), ("max_read_bytes", C.c_uint64), ("max_read_len", C.c_uint32), ("cigar_ops_per_read", C.c_uint32),
                ("n_slots", C.c_uint32), ("fillers_per_device", C.c_int32), ("fast_lookup", C.c_int32), ("extract_tags", C.c_int32)]


_host = None


def load_host_library():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise ClqError(L.E_CUDA, "%s is missing: build it with `make -C clique_b200/csrc` (there is no fallback)" % HOST_LIB)
        L.load_library()  # libclq.so first: libclq_host.so links against it through $ORIGIN
        h = C.CDLL(HOST_LIB)
        h.clqh_align_reads_span.restype = C.c_int32
        h.clqh_align_reads_span.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SpanOptions), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                            C.c_uint64, C.c_void_p] + [C.c_double] * 6 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                                                         C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        _host = h
    return _host


def align_reads_span(devices, refs, read_bytes, read_off, scoring, fixed_ref=None, batch_reads=1 << 18, batch_bytes=0, max_read_len=1 << 16,
                     cigar_ops_per_read=32, n_slots=2, fillers_per_device=2, fast_lookup=True, passes=1):
    """Runs ShardedAligner::align_reads_span `passes` times (the first page-locks the staging buffers) and returns
    (BatchResult in input order, stats dict of the LAST pass)."""
    h = load_host_library()
    devs = np.ascontiguousarray(devices, dtype=np.int32)
    rb = np.frombuffer(b"".join(refs), dtype=np.uint8)
    ro = np.zeros(len(refs) + 1, np.uint64)
    ro[1:] = np.cumsum([len(r) for r in refs], dtype=np.uint64)
    qb = np.ascontiguousarray(read_bytes, dtype=np.uint8)
    qo = np.ascontiguousarray(read_off, dtype=np.uint64)
    n = len(qo) - 1
    fr = None if fixed_ref is None else np.ascontiguousarray(fixed_ref, dtype=np.int32)
    opt = SpanOptions(batch_reads, batch_bytes, max_read_len, cigar_ops_per_read, n_slots, fillers_per_device, int(bool(fast_lookup)), 0)
    res = np.zeros(max(n, 1), _RESULT_DT)
    cap = int(max(1024, n * cigar_ops_per_read))
    pool = np.zeros(cap, np.uint32)
    used, scale = C.c_uint64(), C.c_int32(1)
    stats = np.zeros(8 + 3 * len(devs), np.float64)
    err = C.create_string_buffer(512)
    sc = (scoring.match_score, scoring.mismatch_score, scoring.special_character_score, scoring.gap_open, scoring.gap_extend, scoring.final_gap_multiplier)
    rc = h.clqh_align_reads_span(devs.ctypes.data, len(devs), C.byref(opt), rb.ctypes.data, ro.ctypes.data, len(refs), qb.ctypes.data, qo.ctypes.data, n,
                                 fr.ctypes.data if fr is not None else None, *sc, passes, res.ctypes.data, pool.ctypes.data, cap, C.byref(used),
                                 C.byref(scale), stats.ctypes.data, err, len(err))
    if rc != L.CLQ_OK:
        raise ClqError(rc, "clqh_align_reads_span: %s" % err.value.decode(errors="replace"))
    r = res[:n]
    out = BatchResult(scale.value, r["score_scaled"].copy(), r["ref_index"].copy(), r["cigar_off"].copy(), r["cigar_len"].copy(), r["status"].copy(),
                      pool[:used.value].copy(), matches=r["matches"].copy(), mismatches=r["mismatches"].copy())
    st = {"seconds": float(stats[0]), "setup_seconds": float(stats[1]), "fill_seconds": float(stats[2]), "sink_seconds": float(stats[3]),
          "reads": int(stats[4]), "aligned": int(stats[5]), "dropped": int(stats[6]), "batches": int(stats[7]),
          "device_kernel_ms": [float(stats[8 + 3 * d]) for d in range(len(devs))], "device_reads": [int(stats[9 + 3 * d]) for d in range(len(devs))],
          "device_cells": [int(stats[10 + 3 * d]) for d in range(len(devs))]}
    return out, st
