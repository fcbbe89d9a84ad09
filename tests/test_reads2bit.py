"""2-bit packed read ingestion (include/clq.h: clq_pack2 / clq_upload_packed2 / clq_submit_packed2).

CPU: the host packer against a numpy restatement of the format (codes, word layout, exception list, capacity protocol,
round trip).  GPU: a packed batch gives the oracle's results -- and the ASCII upload's, record for record -- including reads
with N, IUPAC and lower-case bytes (match_mismatch compares raw bytes: alignment/scoring_functions.rs:100-102), ragged and
empty reads, and batch sizes around the 16-base word / 64-base vector / 2048-base warp edges of unpack2_kernel."""
import ctypes as C

import numpy as np
import pytest

from clique_b200 import _lib as L
from clique_b200 import AffineScoring, Aligner, ClqError, PackedReads, Reference, ReferenceManager, pack_reads_2bit
from clique_b200.aligner import pack_reads


def np_pack(b):
    """The format, restated: A C G T -> 0 1 2 3 (anything else 0 + exception), base i in bits 2 (i % 16) of word i // 16."""
    b = np.asarray(b, np.uint8)
    code = np.zeros(256, np.uint8)
    plain = np.zeros(256, bool)
    for k, ch in enumerate(b"ACGT"):
        code[ch] = k
        plain[ch] = True
    n = b.size
    c = np.zeros(((n + 15) // 16) * 16, np.uint32)
    c[:n] = code[b]
    words = (c.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    pos = np.nonzero(~plain[b])[0].astype(np.uint64)
    return words, pos, b[pos.astype(np.int64)]


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 63, 64, 65, 2047, 2048, 2049, 100003])
def test_pack2_matches_format(n):
    rng = np.random.default_rng(n)
    b = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if n:
        k = max(1, n // 50)
        b[rng.integers(0, n, k)] = np.frombuffer(b"NacgtRYn-", np.uint8)[rng.integers(0, 9, k)]
    pk = pack_reads_2bit(b if n else np.zeros(1, np.uint8), n)
    w, pos, byt = np_pack(b)
    assert pk.n_bytes == n
    assert np.array_equal(pk.words[:w.size], w)
    assert np.array_equal(pk.exc_pos, pos) and np.array_equal(pk.exc_byte, byt)
    assert np.array_equal(pk.unpack(), b)


def test_pack2_capacity_protocol():
    lib = L.load_library()
    b = np.frombuffer(b"ACGTNNACGTNACGTAn", np.uint8).copy()
    words = np.zeros(2, np.uint32)
    ne = C.c_uint64()
    pos, byt = np.zeros(2, np.uint64), np.zeros(2, np.uint8)
    rc = lib.clq_pack2(b.ctypes.data, b.size, words.ctypes.data, pos.ctypes.data, byt.ctypes.data, 2, C.byref(ne))
    assert rc == L.E_LIMIT and ne.value == 4                      # the capacity needed; the first two entries are written
    assert list(pos) == [4, 5] and bytes(byt) == b"NN"
    assert np.array_equal(words, np_pack(b)[0])                    # the stream itself is complete
    pos, byt = np.zeros(4, np.uint64), np.zeros(4, np.uint8)
    assert lib.clq_pack2(b.ctypes.data, b.size, words.ctypes.data, pos.ctypes.data, byt.ctypes.data, 4, C.byref(ne)) == L.CLQ_OK
    assert list(pos) == [4, 5, 10, 16] and bytes(byt) == b"NNNn"
    assert lib.clq_pack2(None, 5, words.ctypes.data, None, None, 0, C.byref(ne)) == L.E_INVALID
    assert lib.clq_pack2(b.ctypes.data, 0, None, None, None, 0, C.byref(ne)) == L.CLQ_OK and ne.value == 0
    with pytest.raises(ClqError):
        pack_reads_2bit(b, out_words=np.zeros(1, np.uint32))       # too small a word buffer


def test_packed_reads_is_a_quarter_of_the_bytes():
    b = np.frombuffer(b"ACGT" * 1000, np.uint8)
    pk = pack_reads_2bit(b)
    assert pk.words.nbytes * 4 == b.size and pk.exc_pos.size == 0


# ------------------------------------------------------------------------------------------------ GPU
def _mutate(rng, s, p):
    out = bytearray()
    for ch in s:
        r = rng.random()
        if r < p / 3:
            continue
        if r < 2 * p / 3:
            out.append(b"ACGT"[rng.integers(0, 4)])
        out.append(ch if r > p else b"ACGT"[rng.integers(0, 4)])
    return bytes(out)


def _same(a, b):
    for f in ("score_scaled", "ref_index", "cigar_len", "status", "matches", "mismatches"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for i in range(len(a.status)):
        assert np.array_equal(a.cigar(i), b.cigar(i)), i


@pytest.mark.gpu
@pytest.mark.parametrize("search", ["fixed", "exhaustive", "quick"])
def test_packed_batch_equals_ascii_batch_and_oracle(search):
    import _oracle as O
    from test_gpu_parity import compare
    rng = np.random.default_rng(11 + len(search))
    al = Aligner(device=0, max_reads=1 << 12, max_read_bytes=1 << 22, max_read_len=1 << 14)
    refs = [bytes(b"ACGT"[i] for i in rng.integers(0, 4, n)) for n in (215, 180, 260)]
    if search == "fixed":
        refs = refs[:1]
    reads = []
    for k in range(600):
        r = _mutate(rng, refs[int(rng.integers(0, len(refs)))], 0.06)
        m = k % 6
        if m == 1 and r:
            r = r[:5] + b"N" + r[6:]
        elif m == 2 and len(r) > 40:
            r = r[:20] + r[20:40].lower() + r[40:]      # soft-masked bases differ from the reference's upper case
        elif m == 3 and r:
            r = r[:9] + b"RY" + r[11:]
        reads.append(r)
    reads[7] = b""                                       # empty and ragged reads
    reads[8] = b"A"
    reads[9] = b"n"
    reads.append(bytes(b"ACGT"[i] for i in rng.integers(0, 4, 1500)))   # a multi-stripe read
    sc = (10.0, -9.0, 9.0, -20.0, -2.0, 1.0)
    rm = ReferenceManager([Reference(r, b"r%d" % i) for i, r in enumerate(refs)], 8, 4)
    al.set_references(rm)
    qb, qo = pack_reads(reads)
    fixed = np.zeros(len(reads), np.int32) if search == "fixed" else None
    plain = al.align_batch(qb, qo, AffineScoring(*sc), search, "readlen", fixed_ref=fixed, with_stats=True)
    packed = al.align_batch(qb, qo, AffineScoring(*sc), search, "readlen", fixed_ref=fixed, with_stats=True, packed2=True)
    _same(plain, packed)
    pk = al.pack_reads(qb, int(qo[-1]))
    assert isinstance(pk, PackedReads) and pk.exc_pos.size > 0
    again = al.align_batch(pk, qo, AffineScoring(*sc), search, "readlen", fixed_ref=fixed)   # a pre-packed batch
    _same(plain, again)
    rb, ro = O.pack_seqs(refs)
    want = O.align_batch(rb, ro, qb, qo, sc, search=search, fixed_ref=fixed, band_mode="readlen", kmer=(8, 4), threads=8)
    compare(packed, want, len(reads), "packed2 " + search)
    # the point of the form: a quarter of the read bytes on the wire (+ 9 B per exception)
    n_bytes = int(qo[-1])
    other = packed.stats["h2d_bytes"] - ((n_bytes + 15) // 16) * 4 - 9 * pk.exc_pos.size
    assert other == plain.stats["h2d_bytes"] - n_bytes


@pytest.mark.gpu
@pytest.mark.parametrize("n_bases", [1, 15, 16, 17, 63, 64, 65, 2047, 2048, 2049, 8191, 8193])
def test_unpack_kernel_edges(n_bases):
    """One read of n_bases against itself: the score is n * match only if every base of the stream was expanded in place
    (word, 128-bit vector and warp-chunk edges of unpack2_kernel), and a trailing 'N' exercises patch2_kernel's last byte."""
    rng = np.random.default_rng(n_bases)
    ref = bytes(b"ACGT"[i] for i in rng.integers(0, 4, n_bases))
    al = Aligner(device=0, max_reads=16, max_read_bytes=1 << 20, max_read_len=1 << 16, max_ref_bytes=1 << 20)
    al.set_references(ReferenceManager([Reference(ref, b"self")]), build_kmer_index=False)
    reads = [ref, ref[:-1] + b"N"]
    qb, qo = pack_reads(reads)
    sc = AffineScoring(10.0, -9.0, 9.0, -20.0, -2.0, 1.0)
    fixed = np.zeros(2, np.int32)
    br = al.align_batch(qb, qo, sc, "fixed", "readlen", fixed_ref=fixed, packed2=True)
    pl = al.align_batch(qb, qo, sc, "fixed", "readlen", fixed_ref=fixed)
    _same(pl, br)
    assert int(br.status[0]) == 0 and int(br.score_scaled[0]) == 10 * n_bases * br.scale
    assert (int(br.matches[0]), int(br.mismatches[0])) == (n_bases, 0)


@pytest.mark.gpu
def test_packed_large_batch_grid_stride():
    """26 M bases: more 128-bit vectors than unpack2_kernel's grid has lanes, so the grid-stride loop runs; every read is the
    reference with one substitution at a read-specific column (so a misplaced word shows up as a wrong CIGAR-free score or
    mismatch count), every 7th read ends in 'N'."""
    rng = np.random.default_rng(5)
    L1, n = 215, 120_000
    ref = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, L1)].copy()
    qb = np.tile(ref, n).reshape(n, L1)
    col = rng.integers(1, L1 - 1, n)
    sub = np.frombuffer(b"ACGT", np.uint8)[(np.searchsorted(np.frombuffer(b"ACGT", np.uint8), qb[np.arange(n), col]) + 1) % 4]
    qb[np.arange(n), col] = sub
    qb[::7, L1 - 1] = ord("N")
    qb = qb.reshape(-1)
    qo = (np.arange(n + 1, dtype=np.uint64) * L1)
    al = Aligner(device=0, max_reads=n, max_read_bytes=n * L1 + 64, max_read_len=1 << 12, cigar_ops_per_read=4)
    al.set_references(ReferenceManager([Reference(ref.tobytes(), b"amp")]), build_kmer_index=False)
    sc = AffineScoring(10.0, -9.0, 9.0, -20.0, -2.0, 1.0)
    fixed = np.zeros(n, np.int32)
    pl = al.align_batch(qb, qo, sc, "fixed", "readlen", fixed_ref=fixed)
    pk = al.align_batch(qb, qo, sc, "fixed", "readlen", fixed_ref=fixed, packed2=True)
    _same(pl, pk)
    assert np.all(pk.status == 0)
    plain_reads = np.ones(n, bool)
    plain_reads[::7] = False
    assert np.all(pk.score_scaled[plain_reads] == (10 * (L1 - 1) - 9) * pk.scale)
    assert np.all(pk.mismatches[plain_reads] == 1) and np.all(pk.matches[plain_reads] == L1 - 1)


def test_pack2_round_trip_property():
    """any byte string survives pack -> unpack (the exception list carries whatever the 2-bit alphabet cannot), and the
    packer's fast bodies (AVX2 / SWAR, plain runs of 16-base words) agree with the per-byte restatement at every alignment"""
    from hypothesis import given, settings, strategies as st

    dna = st.lists(st.sampled_from(list(b"ACGT")), min_size=0, max_size=200).map(bytes)
    junk = st.binary(min_size=0, max_size=6)
    pieces = st.lists(st.one_of(dna, dna, junk), min_size=0, max_size=12).map(b"".join)

    @settings(max_examples=300, deadline=None)
    @given(pieces, st.integers(0, 15))
    def check(data, shift):
        buf = np.frombuffer(b"A" * shift + data, np.uint8)[shift:]     # also from unaligned addresses
        pk = pack_reads_2bit(buf if buf.size else np.zeros(1, np.uint8), buf.size)
        w, pos, byt = np_pack(buf)
        assert np.array_equal(pk.words[:w.size], w)
        assert np.array_equal(pk.exc_pos, pos) and np.array_equal(pk.exc_byte, byt)
        assert np.array_equal(pk.unpack(), buf)

    check()


def test_packed_entry_points_reject_a_null_context():
    """no GPU needed: the packed upload validates like clq_upload (CLQ_E_INVALID, no crash)"""
    lib = L.load_library()
    w, off = np.zeros(4, np.uint32), np.zeros(2, np.uint64)
    assert lib.clq_upload_packed2(None, 0, 1, w.ctypes.data, off.ctypes.data, None, None, 0, None) == L.E_INVALID
    assert lib.clq_submit_packed2(None, 0, 1, w.ctypes.data, off.ctypes.data, None, None, 0, None, None, 0, 0.9) == L.E_INVALID
