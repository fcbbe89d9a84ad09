"""Host-side mirror of the reference's alignment interface over the libclq C ABI.

Names, argument meaning and error behaviour follow the reference (file:line relative to rust_cmd/src/):
  AffineScoring                     alignment/scoring_functions.rs:65-113
  AlignmentResult / AlignmentTag    alignment/alignment_matrix.rs:58-120, :694-706
  ReferenceManager / Reference      reference/fasta_reference.rs:41-73, :90-146
  align_two_strings                 alignment_manager.rs:231-273
  align_to_reference_choices        alignment_functions.rs:520-631
  quick_/exhaustive_alignment_search alignment_functions.rs:693-827
  align_reads (the batch loop)      alignment_functions.rs:63-257
All arithmetic happens in the CUDA library; this file only batches, packs and unpacks.
"""
import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import ClqError

OPS = "MID"  # CLQ_OP_M / CLQ_OP_I / CLQ_OP_D


def cigar_to_string(ops) -> str:
    return "".join("%d%s" % (int(o) >> 4, OPS[int(o) & 0xF]) for o in ops)


@dataclass
class AffineScoring:
    """AffineScoring, alignment/scoring_functions.rs:65-73."""
    match_score: float
    mismatch_score: float
    special_character_score: float
    gap_open: float
    gap_extend: float
    final_gap_multiplier: float

    @staticmethod
    def default_dna():  # alignment/scoring_functions.rs:77-86
        return AffineScoring(5.0, -4.0, 4.0, -10.0, -0.5, 0.5)

    @staticmethod
    def align_reads_default():  # alignment_functions.rs:104-111
        return AffineScoring(10.0, -9.0, 9.0, -20.0, -2.0, 1.0)

    @staticmethod
    def merger_default():  # merger.rs:130-139
        return AffineScoring(10.0, -5.0, 8.0, -15.0, -1.0, 0.25)

    def to_int(self) -> L.AffineInt:
        out = L.AffineInt()
        rc = L.load_library().clq_affine_from_f64(self.match_score, self.mismatch_score, self.special_character_score,
                                                  self.gap_open, self.gap_extend, self.final_gap_multiplier, C.byref(out))
        if rc != L.CLQ_OK:
            raise ClqError(rc, "scoring %r has no exact scaled-integer form (or gap_open >= 0)" % (self,))
        return out


@dataclass
class ConvexScoring:
    """ConvexScoring, alignment/scoring_functions.rs:36-53: only a gap *function* in the reference (never called)."""
    match_score: float
    mismatch_score: float
    gap_score: float
    gap_open: float
    gap_extend: float

    def match_mismatch(self, a, b):
        return self.match_score if a == b else self.mismatch_score

    def gap(self, length):  # alignment/scoring_functions.rs:50-52 -- ignores gap_score / gap_extend, gap(0) = -inf
        import math
        return self.gap_open + (math.log10(length) if length > 0 else float("-inf"))


@dataclass
class RustBioScoring:
    """rust_bio_alignment's scoring (alignment_functions.rs:48-61): the closure 1 / -1 with a read 'N' matching anything and
    gap open -5 / extend -1 are hard-coded there (the function ignores its own parameters).  Selects CLQ_RUSTBIO: rust-bio
    `Aligner::global` semantics, PARITY UNPINNED (un-vendored `bio = "*"`)."""
    match_score: int = 1
    mismatch_score: int = -1
    gap_open: int = -5
    gap_extend: int = -1

    def to_int(self) -> L.AffineInt:
        out = L.AffineInt()
        rc = L.load_library().clq_rustbio_scoring(self.match_score, self.mismatch_score, self.gap_open, self.gap_extend, C.byref(out))
        if rc != L.CLQ_OK:
            raise ClqError(rc, "rust-bio scoring %r not representable" % (self,))
        return out


@dataclass
class TwoPieceScoring:
    """Two-piece affine ("convex") gaps as this repository defines them (parity unpinned):
    a gap of length k costs max(o1 + k*e1, o2 + k*e2); integer scores."""
    match: int
    mismatch: int
    special: int
    o1: int
    e1: int
    o2: int
    e2: int

    def to_int(self) -> L.ConvexInt:
        return L.ConvexInt(self.match, self.mismatch, self.special, self.o1, self.e1, self.o2, self.e2, -100000)


@dataclass
class Reference:
    """Reference{sequence, name}, reference/fasta_reference.rs:41-46 (the suffix table is out of scope)."""
    sequence: bytes
    name: bytes


class ReferenceManager:
    """ReferenceManager, reference/fasta_reference.rs:64-73.  `references` keeps insertion order; that ascending
    index is the canonical iteration order (the reference's HashMap order is random per process)."""

    def __init__(self, references: Sequence[Reference], kmer_size: int = 8, kmer_skip: int = 4):
        self.references: List[Reference] = list(references)
        self.reference_name_to_ref = {r.name: i for i, r in enumerate(self.references)}
        self.kmer_size, self.kmer_skip = kmer_size, kmer_skip
        self.longest_ref = max((len(r.sequence) for r in self.references), default=0)

    @staticmethod
    def from_fasta_records(records, kmer_size=8, kmer_skip=4):
        return ReferenceManager([Reference(_b(s), _b(n)) for n, s in records], kmer_size, kmer_skip)

    @staticmethod
    def from_fa_file(path, kmer_size=8, kmer_skip=4):  # reference/fasta_reference.rs:127-130
        recs, name, seq = [], None, []
        with open(path, "rb") as f:
            for ln in f:
                ln = ln.strip()
                if ln.startswith(b">"):
                    if name is not None:
                        recs.append((name, b"".join(seq)))
                    name, seq = ln[1:].split()[0], []
                elif ln:
                    seq.append(ln)
        if name is not None:
            recs.append((name, b"".join(seq)))
        return ReferenceManager.from_fasta_records(recs, kmer_size, kmer_skip)

    def packed(self):
        off = np.zeros(len(self.references) + 1, dtype=np.uint64)
        if self.references:
            off[1:] = np.cumsum([len(r.sequence) for r in self.references], dtype=np.uint64)
        data = np.frombuffer(b"".join(r.sequence for r in self.references), dtype=np.uint8)
        return (data if data.size else np.zeros(1, np.uint8)), off


@dataclass
class AlignmentResult:
    """AlignmentResult, alignment/alignment_matrix.rs:694-706.  Gapped strings and `path` are rebuilt from
    CIGAR + sequences (they are pure functions of them, alignment/alignment_matrix.rs:1019-1065)."""
    reference_name: str
    read_name: str
    reference_aligned: bytes
    read_aligned: bytes
    read_quals: Optional[bytes]
    cigar_string: list  # [(op_char, len)]
    path: list          # [(x, y)]
    score: float
    reference_start: int = 0
    read_start: int = 0
    bounding_box: None = None
    status: int = 0

    @staticmethod
    def from_cigar(ref_name, read_name, reference, read, quals, cigar_ops, score, status=0):
        r, q, x, y = bytearray(), bytearray(), 0, 0
        cig, path = [], []
        n_ops = len(cigar_ops)
        for k, o in enumerate(cigar_ops):
            n, c = int(o) >> 4, int(o) & 0xF
            cig.append((OPS[c], n))
            # the leading boundary run (emitted after the main loop, :1054-1065) is not part of `path`
            boundary = (x == 0 or y == 0) and c != 0 and k == 0 and n_ops > 0
            if c == 0:
                r += reference[x:x + n]; q += read[y:y + n]
                path.extend((x + i + 1, y + i + 1) for i in range(n)); x += n; y += n
            elif c == 2:
                r += reference[x:x + n]; q += b"-" * n
                if not boundary:
                    path.extend((x + i + 1, y) for i in range(n))
                x += n
            else:
                r += b"-" * n; q += read[y:y + n]
                if not boundary:
                    path.extend((x, y + i + 1) for i in range(n))
                y += n
        return AlignmentResult(ref_name, read_name, bytes(r), bytes(q), quals, cig, path, score, 0, 0, None, status)

    def cigar(self) -> str:
        return "".join("%d%s" % (n, c) for c, n in self.cigar_string)


@dataclass
class AlignmentWithRef:
    """AlignmentWithRef, alignment_functions.rs:451-456."""
    alignment: Optional[AlignmentResult]
    ref_name: bytes
    ref_sequence: bytes


@dataclass
class BatchResult:
    """Raw per-read records of one batch (clq_result_t + the CIGAR pool)."""
    scale: int
    score_scaled: np.ndarray
    ref_index: np.ndarray
    cigar_off: np.ndarray
    cigar_len: np.ndarray
    status: np.ndarray
    cigar_pool: np.ndarray
    stats: dict = field(default_factory=dict)
    matches: Optional[np.ndarray] = None      # get_reference_alignment_rate's counters, fused into the traceback walk
    mismatches: Optional[np.ndarray] = None
    tags: Optional[np.ndarray] = None         # [n_reads, tag_stride] read bytes aligned to the tag columns (CLQ_EXTRACT_TAGS)

    def tag_strings(self, i, reference):
        """extract_tagged_sequences' digit keys for read i (extractor.rs:271-332): {symbol byte: read bytes aligned to the
        reference columns holding that symbol}, rebuilt from the GPU's per-column bytes."""
        ref = np.frombuffer(_b(reference), dtype=np.uint8)
        cols = ref[(ref >= 48) & (ref <= 57)]
        row = self.tags[i, :len(cols)]
        return {int(d): bytes(row[cols == d]) for d in np.unique(cols)}

    def alignment_rate(self, i):
        """get_reference_alignment_rate (consensus/consensus_builders.rs:288-307): the `rm` tag of read i (NaN when 0/0)."""
        m, mm = float(self.matches[i]), float(self.mismatches[i])
        return m / (m + mm) if m + mm > 0 else float("nan")

    @property
    def score(self):
        return self.score_scaled.astype(np.float64) / float(self.scale)

    def cigar(self, i):
        o, n = int(self.cigar_off[i]), int(self.cigar_len[i])
        return self.cigar_pool[o:o + n]

    def cigar_string(self, i):
        return cigar_to_string(self.cigar(i))


def get_reference_alignment_rate(reference_aligned, read_aligned):
    """Host mirror of get_reference_alignment_rate (consensus/consensus_builders.rs:288-307) over gapped strings; the batch
    path gets the same counters from the GPU walk (BatchResult.matches / .mismatches)."""
    m = mm = 0
    for a, b in zip(_b(reference_aligned), _b(read_aligned)):
        if a > 64 and a != 78 and b > 64:
            if a == b:
                m += 1
            else:
                mm += 1
    return m / (m + mm) if m + mm else float("nan")


def _b(x):
    return x if isinstance(x, (bytes, bytearray)) else (x.encode() if isinstance(x, str) else bytes(x))


def pack_reads(reads):
    """list of byte strings -> (uint8 array, uint64 offsets)."""
    reads = [_b(r) for r in reads]
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    data = np.frombuffer(b"".join(reads), dtype=np.uint8)
    return (data if data.size else np.zeros(1, np.uint8)), off


@dataclass
class PackedReads:
    """A batch in the 2-bit form of clq_upload_packed2 (include/clq.h): A C G T = 0 1 2 3, base i in bits 2 (i % 16) of
    words[i // 16]; bytes outside that alphabet as (exc_pos, exc_byte), ascending."""
    words: np.ndarray
    exc_pos: np.ndarray
    exc_byte: np.ndarray
    n_bytes: int

    def unpack(self) -> np.ndarray:
        """Host-side inverse (tests / tools): what unpack2_kernel + patch2_kernel leave in the device's read buffer."""
        w = np.ascontiguousarray(self.words[:(self.n_bytes + 15) // 16], dtype=np.uint32)
        codes = (w[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3
        out = np.frombuffer(b"ACGT", np.uint8)[codes.reshape(-1)][:self.n_bytes].copy()
        out[self.exc_pos.astype(np.int64)] = self.exc_byte
        return out


def pack_reads_2bit(read_bytes, n_bytes=None, out_words=None, lib=None) -> PackedReads:
    """clq_pack2 (host only, no GPU needed): raw ASCII batch bytes -> PackedReads."""
    lib = lib or L.load_library()
    rb = np.ascontiguousarray(read_bytes, dtype=np.uint8)
    nb = int(rb.size if n_bytes is None else n_bytes)
    if nb > rb.size:
        raise ClqError(L.E_INVALID, "n_bytes exceeds the buffer")
    nw = (nb + 15) // 16
    words = out_words if out_words is not None else np.zeros(max(nw, 1), np.uint32)
    if words.dtype != np.uint32 or words.size < nw or not words.flags.c_contiguous:
        raise ClqError(L.E_INVALID, "out_words must be a contiguous uint32 array of (n_bytes + 15) // 16 words")
    cap = max(64, nb // 64)
    while True:
        pos, byt, ne = np.zeros(cap, np.uint64), np.zeros(cap, np.uint8), C.c_uint64()
        rc = lib.clq_pack2(rb.ctypes.data, nb, words.ctypes.data, pos.ctypes.data, byt.ctypes.data, cap, C.byref(ne))
        if rc == L.E_LIMIT:   # *n_exc tells the capacity the list needs
            cap = int(ne.value)
            continue
        if rc != L.CLQ_OK:
            raise ClqError(rc, "clq_pack2")
        return PackedReads(words, pos[:ne.value], byt[:ne.value], nb)


_RESULT_DT = np.dtype([("score_scaled", "<i4"), ("ref_index", "<u4"), ("cigar_off", "<u4"), ("cigar_len", "<u4"),
                       ("status", "<u4"), ("matches", "<u4"), ("mismatches", "<u4")])


class Aligner:
    """One clq_ctx on one GPU: reference set + stream slots.  Not thread-safe; use one per thread / device."""

    def __init__(self, device=0, max_reads=1 << 20, max_read_bytes=None, max_read_len=1 << 16, max_refs=4096,
                 max_ref_bytes=1 << 26, cigar_ops_per_read=32, n_slots=2):
        self.lib = L.load_library()
        if self.lib.clq_device_count() <= device:
            raise ClqError(L.E_CUDA, "CUDA device %d not available; libclq has no CPU fallback" % device)
        if max_read_bytes is None:
            max_read_bytes = max_reads * 512
        self.limits = L.Limits(max_reads, max_read_bytes, max_read_len, max_refs, max_ref_bytes,
                               max(1024, max_reads * cigar_ops_per_read), n_slots)
        self.ctx = C.c_void_p()
        rc = self.lib.clq_ctx_create(device, C.byref(self.limits), C.byref(self.ctx))
        if rc != L.CLQ_OK:
            raise ClqError(rc, self.lib.clq_strerror(rc).decode())
        self.device = device
        self.n_slots = self.limits.n_slots
        self.rm: Optional[ReferenceManager] = None
        self._pinned = []
        self._res = [self.alloc_pinned(max_reads, _RESULT_DT) for _ in range(self.n_slots)]
        self._pool = [self.alloc_pinned(int(self.limits.cigar_pool_ops), np.uint32) for _ in range(self.n_slots)]
        self._pending = [None] * self.n_slots

    # ---- lifetime ----
    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.clq_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()
            for p in self._pinned:
                self.lib.clq_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != L.CLQ_OK:
            raise ClqError(rc, "%s: %s" % (self.lib.clq_strerror(rc).decode(), self.lib.clq_ctx_last_error(self.ctx).decode()))

    def alloc_pinned(self, n, dtype):
        """numpy array over page-locked host memory (clq_host_alloc): H2D/D2H from it is true async DMA."""
        dt = np.dtype(dtype)
        nbytes = max(int(n), 1) * dt.itemsize
        p = C.c_void_p()
        rc = self.lib.clq_host_alloc(nbytes, C.byref(p))
        if rc != L.CLQ_OK:
            raise ClqError(rc, "clq_host_alloc(%d)" % nbytes)
        self._pinned.append(p)
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=max(int(n), 1))

    def set_option(self, key, value):
        self._check(self.lib.clq_set_option(self.ctx, key.encode(), int(value)))

    # ---- reference set ----
    def set_references(self, rm: ReferenceManager, build_kmer_index=True):
        data, off = rm.packed()
        self._check(self.lib.clq_refs_set(self.ctx, len(rm.references), data.ctypes.data, off.ctypes.data))
        if build_kmer_index and rm.references:
            self._check(self.lib.clq_kmer_index_set(self.ctx, rm.kmer_size, rm.kmer_skip))
        self.rm = rm

    # ---- raw batch interface ----
    @staticmethod
    def _flags(search, band, score_only):
        f = {"fixed": L.SEARCH_FIXED, "exhaustive": L.SEARCH_EXHAUSTIVE, "quick": L.SEARCH_QUICK}[search]
        if isinstance(band, int):  # explicit bandwidth (perform_affine_alignment_bandwidth's `bandwidth`)
            f |= L.BAND_K | (int(band) << L.BAND_K_SHIFT)
        else:
            f |= {"maxlen": L.BAND_MAXLEN, "readlen": L.BAND_READLEN}[band]
        if score_only:
            f |= L.SCORE_ONLY
        return f

    def pack_reads(self, read_bytes, n_bytes=None, out_words=None) -> "PackedReads":
        """clq_pack2: the batch's bytes as a 2-bit stream + exception list (host only).  `out_words`: a (pinned) uint32 buffer
        to pack into."""
        return pack_reads_2bit(read_bytes, n_bytes, out_words, self.lib)

    def upload(self, slot, read_bytes, read_off, fixed_ref=None, packed2=False):
        """clq_upload; `read_bytes` may be a PackedReads (or raw bytes with packed2=True, packed here): clq_upload_packed2."""
        n = len(read_off) - 1
        fr = None
        if fixed_ref is not None:
            fr = np.ascontiguousarray(fixed_ref, dtype=np.int32)
        ro = np.ascontiguousarray(read_off, dtype=np.uint64)
        frp = fr.ctypes.data if fr is not None else None
        if packed2 and not isinstance(read_bytes, PackedReads):
            read_bytes = self.pack_reads(read_bytes, int(ro[-1]) if n else 0)
        if isinstance(read_bytes, PackedReads):
            pk = read_bytes
            if n and int(ro[-1]) > pk.n_bytes:
                raise ClqError(L.E_INVALID, "read offsets run past the packed stream")
            ne = len(pk.exc_pos)
            self._check(self.lib.clq_upload_packed2(self.ctx, slot, n, pk.words.ctypes.data, ro.ctypes.data,
                                                    pk.exc_pos.ctypes.data if ne else None, pk.exc_byte.ctypes.data if ne else None,
                                                    ne, frp))
            self._pending[slot] = [n, None, (pk, ro, fr)]
            return
        rb = np.ascontiguousarray(read_bytes, dtype=np.uint8)
        self._check(self.lib.clq_upload(self.ctx, slot, n, rb.ctypes.data, ro.ctypes.data, frp))
        self._pending[slot] = [n, None, (rb, ro, fr)]

    def launch(self, slot, scoring, search="fixed", band="readlen", score_only=False, threshold=0.90, extract_tags=False):
        sci = scoring if isinstance(scoring, (L.AffineInt, L.ConvexInt)) else scoring.to_int()
        flags = self._flags(search, band, score_only)
        if isinstance(sci, L.ConvexInt):
            flags |= L.CONVEX
        if isinstance(scoring, RustBioScoring):
            flags |= L.RUSTBIO
        if extract_tags:
            flags |= L.EXTRACT_TAGS
        self._check(self.lib.clq_launch(self.ctx, slot, C.byref(sci), flags, threshold))
        self._pending[slot][1] = getattr(sci, "scale", 1)
        self._pending[slot].append(bool(extract_tags))

    def sync(self, slot):
        self._check(self.lib.clq_sync(self.ctx, slot))

    def stats(self, slot):
        st = L.Stats()
        self._check(self.lib.clq_slot_stats(self.ctx, slot, C.byref(st)))
        return {k: getattr(st, k) for k, _ in L.Stats._fields_}

    def submit(self, slot, read_bytes, read_off, scoring, search="fixed", band="readlen", fixed_ref=None, score_only=False,
               threshold=0.90, extract_tags=False, packed2=False):
        """clq_submit (clq_submit_packed2 for a PackedReads batch): asynchronous H2D + kernels + D2H on the slot's stream."""
        self.upload(slot, read_bytes, read_off, fixed_ref, packed2)
        self.launch(slot, scoring, search, band, score_only, threshold, extract_tags)
        self._check(self.lib.clq_download(self.ctx, slot))

    def wait(self, slot, copy=True, with_stats=False) -> BatchResult:
        n, scale, _keep = self._pending[slot][:3]
        want_tags = len(self._pending[slot]) > 3 and self._pending[slot][-1]
        del self._pending[slot][3:]
        used = C.c_uint64()
        res, pool = self._res[slot], self._pool[slot]
        self._check(self.lib.clq_wait(self.ctx, slot, res.ctypes.data, pool.ctypes.data, len(pool), C.byref(used)))
        r = res[:n]
        cp = (lambda a: a.copy()) if copy else (lambda a: a)
        out = BatchResult(scale, cp(r["score_scaled"]), cp(r["ref_index"]), cp(r["cigar_off"]), cp(r["cigar_len"]),
                          cp(r["status"]), cp(pool[:used.value]), matches=cp(r["matches"]), mismatches=cp(r["mismatches"]))
        if want_tags:
            stride = C.c_uint32()
            self._check(self.lib.clq_tags_download(self.ctx, slot, None, 0, C.byref(stride)))
            buf = np.zeros((n, stride.value), np.uint8)
            if buf.size:
                self._check(self.lib.clq_tags_download(self.ctx, slot, buf.ctypes.data, buf.size, C.byref(stride)))
            out.tags = buf
        if with_stats:
            out.stats = self.stats(slot)
        return out

    def align_batch(self, read_bytes, read_off, scoring, search="fixed", band="readlen", fixed_ref=None, score_only=False,
                    threshold=0.90, with_stats=False, extract_tags=False, packed2=False) -> BatchResult:
        self.submit(0, read_bytes, read_off, scoring, search, band, fixed_ref, score_only, threshold, extract_tags, packed2)
        return self.wait(0, with_stats=with_stats)

    # ---- the reference's call surface ----
    def align_two_strings(self, reference_sequence, read_sequence, read_qual, scoring_function: AffineScoring, local=False,
                          ref_name="ref", read_name="read", _reference_manager=None) -> AlignmentResult:
        """align_two_strings, alignment_manager.rs:231-273 (fresh matrix, bandwidth = max(L1, L2))."""
        if local:
            raise ClqError(L.E_UNSUPPORTED, "local alignment is outside the hot path (SURVEY.md section 2)")
        ref, read = _b(reference_sequence), _b(read_sequence)
        saved = self.rm
        self.set_references(ReferenceManager([Reference(ref, _b(ref_name))]), build_kmer_index=False)
        try:
            rb, ro = pack_reads([read])
            br = self.align_batch(rb, ro, scoring_function, "fixed", "maxlen", fixed_ref=[0])
        finally:  # always restore: an aligner that had no references must not keep the caller's last pairwise reference
            if saved is not None:
                self.set_references(saved)
            else:
                self._check(self.lib.clq_refs_set(self.ctx, 0, None, None))
                self.rm = None
        return self._result(br, 0, ref, read, read_qual, ref_name, read_name)

    def _result(self, br: BatchResult, i, ref, read, quals, ref_name, read_name) -> AlignmentResult:
        st = int(br.status[i])
        if st == L.TRACEBACK_DIVERGED:
            raise ClqError(st, "the reference's traceback does not terminate for this pair (stale band cell)")
        if st != L.CLQ_OK:
            raise ClqError(st, self.lib.clq_strerror(st).decode())
        return AlignmentResult.from_cigar(ref_name, read_name, ref, read, quals, br.cigar(i), float(br.score[i]), st)

    def _search(self, read_name, read, qual, scoring, search, threshold=0.90) -> Optional[AlignmentWithRef]:
        rm = self.rm
        if rm is None or not rm.references:
            return None
        read = _b(read)
        rb, ro = pack_reads([read])
        br = self.align_batch(rb, ro, scoring, search, "readlen", threshold=threshold)
        st = int(br.status[0])
        if st == L.NO_CANDIDATE:
            return None
        if st != L.CLQ_OK:  # ref_index is only meaningful for CLQ_OK records (clique_host.cpp: Aligner::search)
            self._result(br, 0, b"", read, qual, "", read_name)  # raises ClqError with the status
        r = rm.references[int(br.ref_index[0])]
        al = self._result(br, 0, r.sequence, read, qual, r.name.decode(), read_name)
        return AlignmentWithRef(al, r.name, r.sequence)

    def exhaustive_alignment_search(self, read_name, read, qual_sequence, scoring: AffineScoring):
        """exhaustive_alignment_search, alignment_functions.rs:769-827 (ascending index, last maximum wins)."""
        return self._search(read_name, read, qual_sequence, scoring, "exhaustive")

    def quick_alignment_search(self, read_name, read, qual_sequence, scoring: AffineScoring, match_threshold=0.90):
        """quick_alignment_search, alignment_functions.rs:693-767."""
        return self._search(read_name, read, qual_sequence, scoring, "quick", match_threshold)

    def align_to_reference_choices(self, read_name, read, qual_sequence, fast_lookup, scoring: AffineScoring, rust_bio=False):
        """align_to_reference_choices, alignment_functions.rs:520-631.  0 references -> None; > 1 -> quick (fast_lookup) or
        exhaustive search; 1 reference -> with rust_bio=True what the reference does today (:544-603: rust-bio global with the
        hard-coded 1/-1/-5/-1, score reported as 0.0, empty path; PARITY UNPINNED), otherwise clique's own Gotoh with
        bandwidth = read.len() (the call the reference has commented out at :586-597; pinned)."""
        rm = self.rm
        if rm is None or not rm.references:
            return None
        if len(rm.references) == 1:
            read = _b(read)
            rb, ro = pack_reads([read])
            r = rm.references[0]
            if rust_bio:
                br = self.align_batch(rb, ro, RustBioScoring(), "fixed", "maxlen", fixed_ref=[0])
                al = self._result(br, 0, r.sequence, read, qual_sequence, r.name.decode(), read_name)
                al.score, al.path = 0.0, []
                return AlignmentWithRef(al, r.name, r.sequence)
            br = self.align_batch(rb, ro, scoring, "fixed", "readlen", fixed_ref=[0])
            return AlignmentWithRef(self._result(br, 0, r.sequence, read, qual_sequence, r.name.decode(), read_name), r.name, r.sequence)
        return self._search(read_name, read, qual_sequence, scoring, "quick" if fast_lookup else "exhaustive")

    def align_reads(self, reads, scoring: AffineScoring = None, fast_lookup=True, batch_size=None, names=None):
        """The batch loop of align_reads (alignment_functions.rs:135-249) up to the alignment result: drains `reads`
        (byte strings) into batches, keeps every stream slot busy (submit batch k+1 while batch k computes), and yields
        (read_index, AlignmentWithRef or None) in input order."""
        scoring = scoring or AffineScoring.align_reads_default()
        rm = self.rm
        batch_size = batch_size or self.limits.max_reads
        search = "fixed" if len(rm.references) == 1 else ("quick" if fast_lookup else "exhaustive")
        inflight = []  # (slot, base_index, reads)

        def drain(slot, base, chunk):
            br = self.wait(slot)
            for i, rd in enumerate(chunk):
                st = int(br.status[i])
                if st != L.CLQ_OK:  # dropped (too long / no candidate / non-terminating traceback): the reference warns and moves on
                    yield base + i, None
                    continue
                r = rm.references[int(br.ref_index[i])]
                nm = names[base + i] if names else "read%d" % (base + i)
                yield base + i, AlignmentWithRef(AlignmentResult.from_cigar(r.name.decode(), nm, r.sequence, rd, None, br.cigar(i),
                                                                             float(br.score[i])), r.name, r.sequence)

        base, slot, chunk = 0, 0, []
        for rd in reads:
            chunk.append(_b(rd))
            if len(chunk) == batch_size:
                if len(inflight) == self.n_slots:
                    yield from drain(*inflight.pop(0))
                rb, ro = pack_reads(chunk)
                self.submit(slot, rb, ro, scoring, search, "readlen", fixed_ref=np.zeros(len(chunk), np.int32) if search == "fixed" else None)
                inflight.append((slot, base, chunk))
                base += len(chunk); chunk = []; slot = (slot + 1) % self.n_slots
        if chunk:
            if len(inflight) == self.n_slots:
                yield from drain(*inflight.pop(0))
            rb, ro = pack_reads(chunk)
            self.submit(slot, rb, ro, scoring, search, "readlen", fixed_ref=np.zeros(len(chunk), np.int32) if search == "fixed" else None)
            inflight.append((slot, base, chunk))
        for it in inflight:
            yield from drain(*it)


def shard_bounds(read_off, n_shards, ref_len=1, fixed_ref=None):
    """Contiguous read ranges balanced by cells, sum(L1 * L2), not by bytes: on length-sorted input a byte
    split gives the shard of long reads many times the work of the shard of short ones.  `ref_len`: one reference length, or a
    sequence of reference lengths indexed by `fixed_ref` (reads without a usable reference count with the longest).  Returns
    n_shards + 1 boundaries."""
    read_off = np.asarray(read_off, dtype=np.uint64)
    n = len(read_off) - 1
    lens = (read_off[1:] - read_off[:-1]).astype(np.float64)
    rl = np.atleast_1d(np.asarray(ref_len, dtype=np.float64))
    if fixed_ref is not None and rl.size > 1:
        fr = np.asarray(fixed_ref, dtype=np.int64)
        ok = (fr >= 0) & (fr < rl.size)
        l1 = np.where(ok, rl[np.clip(fr, 0, rl.size - 1)], rl.max())
    else:
        l1 = np.full(n, rl.max() if rl.size else 1.0)
    cum = np.concatenate([[0.0], np.cumsum(np.maximum(l1, 1.0) * lens)])
    bounds = [0]
    for k in range(1, n_shards):
        bounds.append(int(np.searchsorted(cum, cum[-1] * k / n_shards, side="left")))
    bounds.append(n)
    for k in range(1, len(bounds)):
        bounds[k] = min(n, max(bounds[k], bounds[k - 1]))
    return bounds


class ShardedAligner:
    """Read-sharded dispatcher over the GPUs of one box: one Aligner (ctx, streams) and one host thread per device,
    contiguous read ranges balanced by cells, no collectives."""

    def __init__(self, devices: Sequence[int], **kw):
        self.aligners = [Aligner(device=d, **kw) for d in devices]

    def set_references(self, rm):
        for a in self.aligners:
            a.set_references(rm)

    def close(self):
        for a in self.aligners:
            a.close()

    def align_batch(self, read_bytes, read_off, scoring, search="fixed", band="readlen", fixed_ref=None, score_only=False) -> BatchResult:
        read_off = np.asarray(read_off, dtype=np.uint64)
        ref_lens = [len(r.sequence) for r in self.aligners[0].rm.references] if self.aligners[0].rm is not None else [1]
        bounds = shard_bounds(read_off, len(self.aligners), ref_lens or [1], fixed_ref)
        outs = [None] * len(self.aligners)
        errs = []

        def work(k):
            try:
                lo, hi = bounds[k], bounds[k + 1]
                off = read_off[lo:hi + 1] - read_off[lo]
                rb = read_bytes[int(read_off[lo]):int(read_off[hi])] if hi > lo else np.zeros(1, np.uint8)
                fr = None if fixed_ref is None else np.asarray(fixed_ref)[lo:hi]
                outs[k] = self.aligners[k].align_batch(rb, off, scoring, search, band, fr, score_only)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=work, args=(k,)) for k in range(len(self.aligners))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return concat_results(outs)


def concat_results(outs: Sequence[BatchResult]) -> BatchResult:
    """Concatenate per-shard results in input order, rebasing CIGAR offsets."""
    base, offs = 0, []
    for o in outs:
        offs.append(o.cigar_off.astype(np.uint32) + np.uint32(base))
        base += len(o.cigar_pool)
    cat = lambda f: np.concatenate([getattr(o, f) for o in outs]) if outs else np.zeros(0)
    have_rate = all(o.matches is not None for o in outs)
    return BatchResult(outs[0].scale, cat("score_scaled"), cat("ref_index"), np.concatenate(offs), cat("cigar_len"),
                       cat("status"), cat("cigar_pool"), matches=cat("matches") if have_rate else None,
                       mismatches=cat("mismatches") if have_rate else None)
