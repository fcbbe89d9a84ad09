"""Property tests of the oracle itself: the scaled-int recurrence equals the f64 one,
band quirks (fact 4), degenerate inputs, threading invariance."""
import numpy as np
import pytest

import _oracle as O

SCORINGS = {
    "cli": (10.0, -9.0, 9.0, -20.0, -2.0, 1.0),            # alignment_functions.rs:104-111
    "default_dna": (5.0, -4.0, 4.0, -10.0, -0.5, 0.5),     # alignment/scoring_functions.rs:77-86
    "merger": (10.0, -5.0, 8.0, -15.0, -1.0, 0.25),        # merger.rs:130-139
    "test": (6.0, -6.0, 5.0, -10.0, -10.0, 1.0),           # alignment/alignment_matrix.rs:1199-1206
}


def rand_seq(rng, n, alphabet=b"ACGTN"):
    return bytes(rng.choice(list(alphabet), size=n).astype(np.uint8)) if n else b""


def mutate(rng, s, p=0.1):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < p / 3:
            continue
        if r < 2 * p / 3:
            out.append(rng.choice(list(b"ACGT")))
            continue
        if r < p:
            out += rand_seq(rng, int(rng.integers(1, 4)), b"ACGT")
        out.append(c)
    return bytes(out)


@pytest.mark.parametrize("name", list(SCORINGS))
def test_int_equals_f64_random(name):
    rng = np.random.default_rng(hash(name) % 2**32)
    sc = SCORINGS[name]
    rc, sci = O.affine_int(sc)
    assert rc == O.OK
    n_div = 0
    for it in range(300):
        l1 = int(rng.integers(0, 60))
        ref = rand_seq(rng, l1, b"ACGTN#acgt" if it % 5 == 0 else b"ACGTN")
        read = mutate(rng, ref, 0.15) if it % 2 else rand_seq(rng, int(rng.integers(0, 60)))
        for band in ("maxlen", "readlen"):
            a = O.align_pair(ref, read, sc, band)
            b = O.align_pair_int(ref, read, sci, band)
            assert b["score_scaled"] == a["score"] * sci.scale, (ref, read, band)
            assert b["status"] == a["status"]
            n_div += a["status"] == O.TRACEBACK_DIVERGED
            if a["status"] == O.OK:
                assert O.cigar_str(b["cigar"]) == O.cigar_str(a["cigar"]), (ref, read, band)
                ra, qa = O.apply_cigar(ref, read, a["cigar"])
                assert ra.replace(b"-", b"") == ref and qa.replace(b"-", b"") == read
    assert n_div < 300


def test_scale_factors():
    assert O.affine_int(SCORINGS["cli"])[1].scale == 1
    assert O.affine_int(SCORINGS["default_dna"])[1].scale == 4
    assert O.affine_int(SCORINGS["merger"])[1].scale == 4
    rc, _ = O.affine_int((1.0, -1.0, 1.0, -1.0 / 3.0, -1.0, 1.0))
    assert rc == O.NOT_REPRESENTABLE
    s = O.affine_int(SCORINGS["default_dna"])[1]
    assert (s.match, s.mismatch, s.special, s.oe_in, s.e_in, s.oe_fin, s.e_fin, s.b0, s.b1, s.max_neg) == \
           (20, -16, 16, -42, -2, -41, -1, -20, -1, -400000)


def test_empty_inputs():
    sc = SCORINGS["cli"]
    r = O.align_pair(b"", b"", sc)
    assert r["score"] == 0.0 and len(r["cigar"]) == 0
    r = O.align_pair(b"ACGT", b"", sc)
    assert r["score"] == -28.0 and O.cigar_str(r["cigar"]) == "4D"
    r = O.align_pair(b"", b"ACG", sc)
    assert r["score"] == -26.0 and O.cigar_str(r["cigar"]) == "3I"


def test_f64_band_quirk():
    # (1/49)*49 truncates to 0 => row 1 skips the last column even for equal lengths
    lo, hi = np.zeros(1, np.int64), np.zeros(1, np.int64)
    import ctypes as C
    a, b = C.c_int64(), C.c_int64()
    O.lib().orc_band(1, 48, 48, 48, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (1, 48)          # y = 48 skipped
    O.lib().orc_band(1, 47, 47, 47, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (1, 48)          # y = 47 included
    # a read much shorter than the reference leaves stale cells that can capture the traceback
    rng = np.random.default_rng(7)
    seen = set()
    for _ in range(200):
        ref = rand_seq(rng, int(rng.integers(20, 80)), b"ACGT")
        read = rand_seq(rng, int(rng.integers(1, 12)), b"ACGT")
        seen.add(O.align_pair(ref, read, SCORINGS["cli"], "readlen")["status"])
    assert O.TRACEBACK_DIVERGED in seen


def test_batch_threads_invariant(goldens):
    rng = np.random.default_rng(3)
    recs = goldens["fastas"]["18guide1_pcr_sequence.first64"][:6]
    refs = [r["seq"].encode() for r in recs]
    reads = [mutate(rng, refs[int(rng.integers(0, len(refs)))], 0.02) for _ in range(24)]
    rb, ro = O.pack_seqs(refs)
    qb, qo = O.pack_seqs(reads)
    outs = [O.align_batch(rb, ro, qb, qo, SCORINGS["cli"], search=s, threads=t, traceback_all=tb)
            for (s, t, tb) in (("exhaustive", 1, True), ("exhaustive", 4, False), ("quick", 3, True))]
    for o in outs[1:]:
        assert (o["score"] == outs[0]["score"]).all() and (o["ref_index"] == outs[0]["ref_index"]).all()
        assert (o["cigar_len"] == outs[0]["cigar_len"]).all() and (o["cigar_pool"] == outs[0]["cigar_pool"]).all()
    assert outs[0]["cells"] == sum(len(r) for r in refs) * sum(len(q) for q in reads)
    assert outs[2]["cells"] < outs[0]["cells"]      # the k-mer vote prunes candidates
