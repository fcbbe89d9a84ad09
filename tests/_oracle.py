"""ctypes binding of the CPU oracle (oracle/libclq_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libclq_oracle.so")

OK, READ_TOO_LONG, NOT_REPRESENTABLE, TRACEBACK_DIVERGED, CIGAR_POOL_FULL, NO_CANDIDATE = range(6)
BAND = {"maxlen": 0, "readlen": 1, "k": 2}
SEARCH = {"fixed": 0, "exhaustive": 1, "quick": 2}
OPS = "MID"


class Affine(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("match_score", "mismatch_score", "special_character_score", "gap_open",
                                           "gap_extend", "final_gap_multiplier")]


class AffineInt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("scale", "match", "mismatch", "special", "oe_in", "e_in", "oe_fin", "e_fin",
                                          "b0", "b1", "max_neg")]


class Convex(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("match", "mismatch", "special", "o1", "e1", "o2", "e2", "max_neg")]


class Result(C.Structure):
    _fields_ = [("score", C.c_double), ("status", C.c_int32), ("n_cigar", C.c_uint32), ("aligned_len", C.c_uint32),
                ("path_len", C.c_uint32), ("matches", C.c_uint32), ("mismatches", C.c_uint32)]


class Batch(C.Structure):
    _fields_ = [("n_refs", C.c_uint32), ("ref_bytes", C.c_void_p), ("ref_off", C.c_void_p), ("n_reads", C.c_uint32),
                ("read_bytes", C.c_void_p), ("read_off", C.c_void_p), ("fixed_ref", C.c_void_p),
                ("search_mode", C.c_int32), ("band_mode", C.c_int32), ("band_k", C.c_uint64), ("kmer_k", C.c_uint32),
                ("kmer_skip", C.c_uint32), ("match_threshold", C.c_double), ("threads", C.c_int32),
                ("traceback_all_candidates", C.c_int32)]


class BatchOut(C.Structure):
    _fields_ = [("score", C.c_void_p), ("ref_index", C.c_void_p), ("status", C.c_void_p), ("cigar_off", C.c_void_p),
                ("cigar_len", C.c_void_p), ("cigar_pool", C.c_void_p), ("cigar_cap", C.c_uint64),
                ("cigar_used", C.c_uint64), ("cells", C.c_uint64), ("matches", C.c_void_p), ("mismatches", C.c_void_p)]


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("clq_oracle.c", "clq_oracle.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.run(["make", "-C", ORACLE_DIR, "-B", "libclq_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_match_mismatch.restype = C.c_double
        L.orc_match_mismatch.argtypes = [C.POINTER(Affine), C.c_uint8, C.c_uint8]
        L.orc_three_way_max.restype = C.c_double
        L.orc_three_way_max.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_convex_gap.restype = C.c_double
        L.orc_convex_gap.argtypes = [C.c_double, C.c_size_t]
        L.orc_matrix_create.restype = C.c_void_p
        L.orc_matrix_create.argtypes = [C.c_size_t, C.c_size_t]
        L.orc_matrix_free.argtypes = [C.c_void_p]
        L.orc_fill.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(Affine), C.c_size_t]
        L.orc_traceback.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(Result),
                                    C.c_void_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_alignment_rate.restype = C.c_double
        L.orc_alignment_rate.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_simplify_cigar.restype = C.c_size_t
        L.orc_simplify_cigar.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_affine_to_int.argtypes = [C.POINTER(Affine), C.POINTER(AffineInt)]
        L.orci_align_pair.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(AffineInt), C.c_int,
                                      C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                      C.c_void_p, C.c_size_t, C.c_int]
        L.orc_convex_align_pair.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(Convex),
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                            C.c_void_p, C.c_size_t, C.c_int]
        L.orc_band.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_kmer_index_build.restype = C.c_void_p
        L.orc_kmer_index_build.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_kmer_index_free.argtypes = [C.c_void_p]
        L.orc_kmer_index_size.restype = C.c_uint32
        L.orc_kmer_index_size.argtypes = [C.c_void_p]
        L.orc_kmer_votes.restype = C.c_uint32
        L.orc_kmer_votes.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_align_batch.argtypes = [C.POINTER(Batch), C.POINTER(Affine), C.POINTER(BatchOut)]
        L.orc_rustbio_global.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t]
        L.orc_phred_to_prob.restype = C.c_double
        L.orc_phred_to_prob.argtypes = [C.c_uint8]
        L.orc_prob_to_phred.restype = C.c_uint8
        L.orc_prob_to_phred.argtypes = [C.c_double]
        L.orc_combine_phred_scores.restype = C.c_uint8
        L.orc_combine_phred_scores.argtypes = [C.c_uint8, C.c_uint8, C.c_int]
        L.orc_alignment_rate_and_consensus.restype = C.c_size_t
        L.orc_alignment_rate_and_consensus.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t,
                                                       C.c_void_p, C.c_void_p]
        L.orc_find_greedy_non_overlapping_segments.restype = C.c_size_t
        L.orc_find_greedy_non_overlapping_segments.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                               C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_orient_by_longest_segment.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t),
                                                    C.POINTER(C.c_size_t)]
        L.orc_extract_tagged_sequences.restype = C.c_size_t
        L.orc_extract_tagged_sequences.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_reverse_complement.restype = None
        L.orc_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def affine(d):
    """dict / tuple -> Affine"""
    if isinstance(d, Affine):
        return d
    if isinstance(d, dict):
        return Affine(d["match_score"], d["mismatch_score"], d["special_character_score"], d["gap_open"],
                      d["gap_extend"], d["final_gap_multiplier"])
    return Affine(*d)


def affine_int(sc):
    out = AffineInt()
    rc = lib().orc_affine_to_int(C.byref(affine(sc)), C.byref(out))
    return rc, out


def cigar_str(ops):
    return "".join("%d%s" % (int(o) >> 4, OPS[int(o) & 0xF]) for o in ops)


def cigar_parse(s):
    import re
    return np.array([(int(n) << 4) | OPS.index(c) for n, c in re.findall(r"(\d+)([MID])", s)], dtype=np.uint32)


def bandwidth(mode, l1, l2, k=0):
    return {"maxlen": max(l1, l2), "readlen": l2, "k": k}[mode]


def align_pair(ref, read, sc, band_mode="maxlen", band_k=0, dim=None):
    """f64 oracle: fill + traceback on a fresh matrix.  Returns dict(score, status, cigar, ref_aligned, read_aligned)."""
    L = lib()
    ref, read = bytes(ref), bytes(read)
    l1, l2 = len(ref), len(read)
    a, b = dim if dim else (l1 + 1, l2 + 1)
    m = L.orc_matrix_create(a, b)
    try:
        s = affine(sc)
        rc = L.orc_fill(m, ref, l1, read, l2, C.byref(s), bandwidth(band_mode, l1, l2, band_k))
        assert rc == 0
        res = Result()
        cig = np.zeros(l1 + l2 + 2, dtype=np.uint32)
        a1 = C.create_string_buffer(l1 + l2 + 2)
        a2 = C.create_string_buffer(l1 + l2 + 2)
        L.orc_traceback(m, ref, l1, read, l2, C.byref(res), cig.ctypes.data, len(cig), a1, a2, l1 + l2 + 2)
        return {"score": res.score, "status": res.status, "cigar": cig[:res.n_cigar].copy(),
                "ref_aligned": a1.raw[:res.aligned_len], "read_aligned": a2.raw[:res.aligned_len],
                "path_len": res.path_len, "matches": res.matches, "mismatches": res.mismatches}
    finally:
        L.orc_matrix_free(m)


def align_pair_int(ref, read, sci, band_mode="maxlen", band_k=0, traceback=True):
    L = lib()
    ref, read = bytes(ref), bytes(read)
    l1, l2 = len(ref), len(read)
    score, status, n = C.c_int64(), C.c_int32(), C.c_uint32()
    cig = np.zeros(l1 + l2 + 2, dtype=np.uint32)
    L.orci_align_pair(ref, l1, read, l2, C.byref(sci), BAND[band_mode], band_k, C.byref(score), C.byref(status),
                      C.byref(n), cig.ctypes.data, len(cig), 1 if traceback else 0)
    return {"score_scaled": score.value, "status": status.value, "cigar": cig[:n.value].copy()}


def convex_align_pair(ref, read, cv, traceback=True):
    L = lib()
    ref, read = bytes(ref), bytes(read)
    l1, l2 = len(ref), len(read)
    score, status, n = C.c_int64(), C.c_int32(), C.c_uint32()
    cig = np.zeros(l1 + l2 + 2, dtype=np.uint32)
    L.orc_convex_align_pair(ref, l1, read, l2, C.byref(cv), C.byref(score), C.byref(status), C.byref(n),
                            cig.ctypes.data, len(cig), 1 if traceback else 0)
    return {"score": score.value, "status": status.value, "cigar": cig[:n.value].copy()}


def pack_seqs(seqs):
    """list of bytes -> (uint8 array, uint64 offsets)"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    data = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    if data.size == 0:
        data = np.zeros(1, np.uint8)
    return data, off


def align_batch(ref_bytes, ref_off, read_bytes, read_off, sc, search="fixed", fixed_ref=None, band_mode="readlen",
                band_k=0, kmer=(8, 4), threshold=0.90, threads=1, traceback_all=True, cigar_cap=None):
    """Batch oracle over packed arrays.  Returns dict of numpy arrays."""
    L = lib()
    n_refs, n_reads = len(ref_off) - 1, len(read_off) - 1
    ref_bytes = np.ascontiguousarray(ref_bytes, dtype=np.uint8)
    read_bytes = np.ascontiguousarray(read_bytes, dtype=np.uint8)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
    read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    if cigar_cap is None:
        cigar_cap = int(n_reads) * 64 + int(read_off[-1]) + 1024
    out = {"score": np.zeros(n_reads, np.float64), "ref_index": np.zeros(n_reads, np.uint32),
           "status": np.zeros(n_reads, np.uint32), "cigar_off": np.zeros(n_reads, np.uint64),
           "cigar_len": np.zeros(n_reads, np.uint32), "cigar_pool": np.zeros(cigar_cap, np.uint32),
           "matches": np.zeros(n_reads, np.uint32), "mismatches": np.zeros(n_reads, np.uint32)}
    fr = None
    if fixed_ref is not None:
        fr = np.ascontiguousarray(fixed_ref, dtype=np.int32)
    b = Batch(n_refs, ref_bytes.ctypes.data, ref_off.ctypes.data, n_reads, read_bytes.ctypes.data, read_off.ctypes.data,
              fr.ctypes.data if fr is not None else None, SEARCH[search], BAND[band_mode], band_k, kmer[0], kmer[1],
              threshold, threads, 1 if traceback_all else 0)
    o = BatchOut(out["score"].ctypes.data, out["ref_index"].ctypes.data, out["status"].ctypes.data,
                 out["cigar_off"].ctypes.data, out["cigar_len"].ctypes.data, out["cigar_pool"].ctypes.data, cigar_cap, 0, 0,
                 out["matches"].ctypes.data, out["mismatches"].ctypes.data)
    s = affine(sc)
    rc = L.orc_align_batch(C.byref(b), C.byref(s), C.byref(o))
    out["rc"] = rc
    out["cigar_used"] = o.cigar_used
    out["cells"] = o.cells
    out["cigar_pool"] = out["cigar_pool"][:o.cigar_used]
    return out


def alignment_rate(ref_aligned, read_aligned):
    m, mm = C.c_uint32(), C.c_uint32()
    r = lib().orc_alignment_rate(bytes(ref_aligned), bytes(read_aligned), len(ref_aligned), C.byref(m), C.byref(mm))
    return r, m.value, mm.value


def apply_cigar(ref, read, cigar):
    """gapped strings from CIGAR + sequences (what the host layer rebuilds)"""
    r, q, x, y = bytearray(), bytearray(), 0, 0
    for o in cigar:
        n, c = int(o) >> 4, int(o) & 0xF
        if c == 0:
            r += ref[x:x + n]; q += read[y:y + n]; x += n; y += n
        elif c == 2:
            r += ref[x:x + n]; q += b"-" * n; x += n
        else:
            r += b"-" * n; q += read[y:y + n]; y += n
    return bytes(r), bytes(q)


def parse_tag_records(buf):
    """[key u8][len u32 LE][bytes] records -> {key: bytes}"""
    out, i = {}, 0
    while i < len(buf):
        k, n = buf[i], int.from_bytes(buf[i + 1:i + 5], "little")
        out[k] = bytes(buf[i + 5:i + 5 + n])
        i += 5 + n
    return out


def extract_tagged_sequences(aligned_read, aligned_ref):
    """extractor.rs:271-332 (zips the two strings: stops at the shorter) -> {key byte: bytes}"""
    n = min(len(aligned_read), len(aligned_ref))
    cap = 16 * 256 + 3 * n + 64
    buf = C.create_string_buffer(cap)
    w = lib().orc_extract_tagged_sequences(bytes(aligned_read), bytes(aligned_ref), n, buf, cap)
    return parse_tag_records(buf.raw[:w])


def reverse_complement(dna):
    buf = C.create_string_buffer(max(1, len(dna)))
    lib().orc_reverse_complement(bytes(dna), len(dna), buf)
    return buf.raw[:len(dna)]


RUSTBIO_CLI = (1, -1, -5, -1)  # rust_bio_alignment's hard-coded scoring, alignment_functions.rs:55-57


def rustbio_global(ref, read, scoring=RUSTBIO_CLI):
    """rust-bio Aligner::global as the single-reference branch calls it (PARITY UNPINNED restatement, oracle/clq_oracle.h)."""
    ref, read = bytes(ref), bytes(read)
    score, n = C.c_int32(), C.c_uint32()
    cig = np.zeros(len(ref) + len(read) + 2, dtype=np.uint32)
    rc = lib().orc_rustbio_global(ref, len(ref), read, len(read), scoring[0], scoring[1], scoring[2], scoring[3],
                                  C.byref(score), C.byref(n), cig.ctypes.data, len(cig))
    return {"score": score.value, "status": rc, "cigar": cig[:n.value].copy()}


def alignment_rate_and_consensus(a1, q1, a2, q2):
    """merger.rs:428-498 -> (bases, quals), or None where the reference panics (quality index out of bounds)"""
    n = len(a1)
    assert len(a2) == n
    ob, oq = C.create_string_buffer(max(1, n)), C.create_string_buffer(max(1, n))
    r = lib().orc_alignment_rate_and_consensus(bytes(a1), bytes(q1), len(q1), bytes(a2), bytes(q2), len(q2), n, ob, oq)
    if r == 2 ** 64 - 1:
        return None
    return ob.raw[:n], oq.raw[:n]


def merge_reads_by_alignment(read1, qual1, read2, qual2, sc):
    """merge_reads_by_alignment, merger.rs:348-396: align(read1, revcomp(read2)) full matrix, then the consensus"""
    rc2, q2r = reverse_complement(read2), bytes(qual2)[::-1]
    a = align_pair(read1, rc2, sc, "maxlen")
    if a["status"] != OK:
        return None
    return alignment_rate_and_consensus(a["ref_aligned"], qual1, a["read_aligned"], q2r)


def find_greedy_non_overlapping_segments(search, reference, seed_size):
    """linked_alignment.rs:97-130 -> ([(search_start, ref_start, length)], start_position)"""
    search, reference = bytes(search), bytes(reference)
    cap = len(search) + 2
    seg = np.zeros((cap, 3), np.uint32)
    sp = C.c_size_t()
    n = lib().orc_find_greedy_non_overlapping_segments(search, len(search), reference, len(reference), seed_size, seg.ctypes.data, cap, C.byref(sp))
    return [tuple(int(v) for v in seg[i]) for i in range(n)], sp.value


def orient_by_longest_segment(search, reference, seed_size):
    """linked_alignment.rs:24-32 -> (forward?, fwd_score, rev_score)"""
    f, r = C.c_size_t(), C.c_size_t()
    fwd = lib().orc_orient_by_longest_segment(bytes(search), len(search), bytes(reference), len(reference), seed_size, C.byref(f), C.byref(r))
    return bool(fwd), f.value, r.value
