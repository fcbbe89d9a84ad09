"""The C++ host layer (include/clique_host.hpp, libclq_host.so): its pure string functions against the oracle and the
reference's golden vectors (CPU, through the extern "C" view include/clqh.h), and -- on the GPU -- the batch loop
ShardedAligner::align_reads through the clq_align driver against the oracle + the Python host layer."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "clique_b200", "libclq_host.so")
CLQ_ALIGN = os.path.join(ROOT, "clique_b200", "clq_align")

SYMBOLS = ["clqh_extract_tagged_sequences", "clqh_reverse_complement", "clqh_f64_to_string", "clqh_get_reference_alignment_rate",
           "clqh_simplify_cigar", "clqh_from_cigar", "clqh_sam_line", "clqh_merge_reads_by_concatenation",
           "clqh_combine_phred_scores", "clqh_alignment_rate_and_consensus", "clqh_merge_read_pairs_by_alignment",
           "clqh_find_greedy_non_overlapping_segments", "clqh_orient_by_longest_segment", "clqh_bam_file", "clqh_extend_hit",
           "clqh_align_reads_span", "clqh_span_claims"]


@pytest.fixture(scope="module")
def H():
    assert os.path.exists(HOST_LIB), "build with `make -C clique_b200/csrc` (python __graft_entry__.py)"
    L = C.CDLL(HOST_LIB)
    L.clqh_extract_tagged_sequences.restype = C.c_size_t
    L.clqh_extract_tagged_sequences.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.clqh_reverse_complement.restype = None
    L.clqh_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    L.clqh_f64_to_string.restype = C.c_size_t
    L.clqh_f64_to_string.argtypes = [C.c_double, C.c_void_p, C.c_size_t]
    L.clqh_get_reference_alignment_rate.restype = C.c_double
    L.clqh_get_reference_alignment_rate.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.clqh_simplify_cigar.restype = C.c_size_t
    L.clqh_simplify_cigar.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.clqh_from_cigar.restype = C.c_int32
    L.clqh_from_cigar.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.clqh_sam_line.restype = C.c_size_t
    L.clqh_sam_line.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.c_double, C.c_int32, C.c_char_p, C.c_void_p, C.c_size_t]
    L.clqh_combine_phred_scores.restype = C.c_uint8
    L.clqh_combine_phred_scores.argtypes = [C.c_uint8, C.c_uint8, C.c_int32]
    L.clqh_alignment_rate_and_consensus.restype = C.c_size_t
    L.clqh_alignment_rate_and_consensus.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t,
                                                    C.c_void_p, C.c_void_p]
    L.clqh_merge_read_pairs_by_alignment.restype = C.c_int32
    L.clqh_merge_read_pairs_by_alignment.argtypes = [C.c_int32, C.c_uint32] + [C.c_void_p] * 6 + [C.c_double] * 6 + [C.c_void_p, C.c_void_p,
                                                                                                                  C.c_uint64, C.c_void_p]
    L.clqh_find_greedy_non_overlapping_segments.restype = C.c_size_t
    L.clqh_find_greedy_non_overlapping_segments.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                                            C.POINTER(C.c_size_t)]
    L.clqh_orient_by_longest_segment.restype = C.c_int32
    L.clqh_orient_by_longest_segment.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t]
    L.clqh_extend_hit.restype = C.c_size_t
    L.clqh_extend_hit.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t]
    L.clqh_bam_file.restype = C.c_size_t
    L.clqh_bam_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_double,
                                C.c_char_p, C.c_void_p, C.c_size_t]
    L.clqh_merge_reads_by_concatenation.restype = C.c_size_t
    L.clqh_merge_reads_by_concatenation.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_size_t]
    return L


def h_extract(H, read, ref):
    cap = 16 * 256 + 3 * max(len(read), len(ref)) + 64
    buf = C.create_string_buffer(cap)
    w = H.clqh_extract_tagged_sequences(read, len(read), ref, len(ref), buf, cap)
    return O.parse_tag_records(buf.raw[:w])


def h_f64(H, v):
    buf = C.create_string_buffer(512)
    n = H.clqh_f64_to_string(v, buf, 512)
    return buf.raw[:n].decode()


def h_from_cigar(H, ref, read, ops):
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    cap = len(ref) + len(read) + 8
    ra, qa = C.create_string_buffer(cap), C.create_string_buffer(cap)
    path = np.zeros(2 * cap, np.uint32)
    al, pl = C.c_size_t(), C.c_size_t()
    rc = H.clqh_from_cigar(ref, len(ref), read, len(read), ops.ctypes.data, len(ops), ra, qa, cap, C.byref(al), path.ctypes.data, cap,
                           C.byref(pl))
    assert rc == 0, rc
    return ra.raw[:al.value], qa.raw[:al.value], path[:2 * pl.value].reshape(-1, 2)


# ------------------------------------------------------------------------------------------------ CPU: pure host functions
def test_host_library_exports(H):
    out = subprocess.run(["nm", "-D", "--defined-only", HOST_LIB], capture_output=True, text=True, check=True).stdout
    names = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for s in SYMBOLS:
        assert s in names, s
    # every symbol include/clqh.h declares is exported
    decl = set(re.findall(r"\b(clqh_\w+)\s*\(", open(os.path.join(ROOT, "include", "clqh.h")).read()))
    assert decl == set(SYMBOLS)


def test_extract_tagged_sequences_goldens_cpp(H, goldens):
    for t in goldens["tagged_sequences"]:
        got = h_extract(H, t["read"].encode(), t["ref"].encode())
        for k, v in t.get("expect", {}).items():
            assert got[ord(k)].decode() == v, (t["name"], k)
        for k in t.get("expect_keys", []):
            assert ord(k) in got
        assert got == O.extract_tagged_sequences(t["read"].encode(), t["ref"].encode()), t["name"]


def test_extract_tagged_sequences_random_vs_oracle(H):
    rng = np.random.default_rng(7)
    alpha = np.frombuffer(b"ACGTacgtNn-0123456789#", dtype=np.uint8)
    for _ in range(300):
        n = int(rng.integers(0, 120))
        # runs, so that extractor regions and tag runs of realistic shape appear
        ref = np.repeat(alpha[rng.integers(0, len(alpha), size=n)], rng.integers(1, 5, size=n))[:n].tobytes()
        read = alpha[rng.integers(0, 9, size=len(ref))].tobytes()
        assert h_extract(H, read, ref) == O.extract_tagged_sequences(read, ref)


def test_reverse_complement_cpp(H, goldens):
    for t in goldens["reverse_complement"]:
        buf = C.create_string_buffer(max(1, len(t["in"])))
        H.clqh_reverse_complement(t["in"].encode(), len(t["in"]), buf)
        assert buf.raw[:len(t["in"])].decode() == t["out"], t
    rng = np.random.default_rng(3)
    for _ in range(100):
        s = rng.integers(33, 127, size=int(rng.integers(0, 80))).astype(np.uint8).tobytes()
        buf = C.create_string_buffer(max(1, len(s)))
        H.clqh_reverse_complement(s, len(s), buf)
        assert buf.raw[:len(s)] == O.reverse_complement(s)


def test_alignment_rate_cpp(H, goldens):
    for t in goldens["alignment_rate"]:
        assert H.clqh_get_reference_alignment_rate(t["ref"].encode(), t["read"].encode(), len(t["ref"])) == t["rate"]
    assert np.isnan(H.clqh_get_reference_alignment_rate(b"NNNN", b"ACGT", 4))


def test_simplify_cigar_cpp(H, goldens):
    for t in goldens["simplify_cigar"]:
        ops = np.array(O.cigar_parse(t["in"]), dtype=np.uint32)
        out = np.zeros(max(1, len(ops)), np.uint32)
        n = H.clqh_simplify_cigar(ops.ctypes.data if len(ops) else None, len(ops), out.ctypes.data)
        assert O.cigar_str(out[:n]) == t["out"], t


# cigar_to_alignment (alignment_functions.rs:830-872): unit Match / Subst / Del / Ins ops -> gapped strings + merged tags.  Here
# that is simplify_cigar_string + AlignmentResult::from_cigar; the reference's four unit tests (:1149-1214), op lists as CIGAR
# letters (Subst is a MatchMismatch step).
CIGAR_TO_ALIGNMENT = [
    (b"ACGT", b"ACGT", "MMMM", b"ACGT", b"ACGT", "4M"),          # :1150-1163
    (b"ACGT", b"AT", "MDDM", b"ACGT", b"A--T", "1M2D1M"),        # :1165-1181 (asserts ref_aln, read_aln[0], read_aln[3], 3 tags)
    (b"AT", b"ACGT", "MIIM", b"A--T", b"ACGT", "1M2I1M"),        # :1183-1199
    (b"ACGT", b"ATGT", "MMMM", b"ACGT", b"ATGT", "4M"),          # :1201-1214 (Match, Subst, Match, Match)
]


def test_cigar_to_alignment_reference_vectors(H):
    for ref, read, unit, want_ref, want_read, want_cigar in CIGAR_TO_ALIGNMENT:
        ops = np.array([(1 << 4) | "MID".index(c) for c in unit], dtype=np.uint32)
        out = np.zeros(len(ops), np.uint32)
        n = H.clqh_simplify_cigar(ops.ctypes.data, len(ops), out.ctypes.data)
        assert O.cigar_str(out[:n]) == want_cigar
        ra, qa, _ = h_from_cigar(H, ref, read, out[:n])
        assert (ra, qa) == (want_ref, want_read)
        assert O.apply_cigar(ref, read, out[:n]) == (want_ref, want_read)       # the oracle-side helper the GPU tests use
        assert O.apply_cigar(ref, read, ops) == (want_ref, want_read)           # unit ops give the same strings


def test_f64_display(H):
    # Rust `Display` for f64 (score.to_string() / rate.to_string() in the BAM tags)
    for v, s in [(979.0, "979"), (391.5, "391.5"), (13.75, "13.75"), (-40.0, "-40"), (0.0, "0"), (1.0, "1"), (0.5, "0.5"),
                 (0.9866666666666667, "0.9866666666666667"), (1e-5, "0.00001"), (2.5e-7, "0.00000025"), (1e21, "1000000000000000000000"),
                 (float("nan"), "NaN"), (float("inf"), "inf"), (float("-inf"), "-inf"), (1 / 3, "0.3333333333333333")]:
        assert h_f64(H, v) == s, (v, h_f64(H, v))
    rng = np.random.default_rng(11)
    for _ in range(500):
        m, mm = int(rng.integers(0, 400)), int(rng.integers(1, 50))
        v = m / (m + mm)
        s = h_f64(H, v)
        assert float(s) == v and "e" not in s and (repr(v) == s or repr(v) == s + ".0" or "e" in repr(v))


def test_from_cigar_matches_oracle_traceback(H, goldens):
    # gapped strings and path rebuilt from (CIGAR, sequences) == what the traceback itself produced
    for p in goldens["pairs"]:
        ref, read = p["ref"].encode(), p["read"].encode()
        a = O.align_pair(ref, read, p["scoring"], p["band_mode"])
        ra, qa, path = h_from_cigar(H, ref, read, a["cigar"])
        assert ra == a["ref_aligned"] and qa == a["read_aligned"], p["name"]
        assert len(path) == a["path_len"], p["name"]  # one entry per unit op of the main loop (leading boundary run excluded)
        assert all(0 <= x <= len(ref) and 0 <= y <= len(read) for x, y in path.tolist())
        e = p["expect"]
        if "ref_aligned" in e:
            assert ra.decode() == e["ref_aligned"] and qa.decode() == e["read_aligned"], p["name"]


def test_sam_line(H, goldens):
    p = goldens["pairs"][2]  # affine_alignment_test: AAAA vs AATAA -> 2M1I2M, score 4
    ref, read = p["ref"].encode(), p["read"].encode()
    a = O.align_pair(ref, read, p["scoring"], "maxlen")
    ops = np.ascontiguousarray(a["cigar"], dtype=np.uint32)
    buf = C.create_string_buffer(4096)
    n = H.clqh_sam_line(b"amp", b"r1", ref, len(ref), read, len(read), ops.ctypes.data, len(ops), a["score"], 0, b"e1=ACGT;rc=1;ar=r1", buf, 4096)
    f = buf.raw[:n].decode().split("\t")
    # to_sam_record, alignment/alignment_matrix.rs:741-771: empty flags, POS = reference_start + 1, SEQ = read without gaps,
    # every quality forced to raw 'H' (72 -> 'i' in SAM text), tags rm / rs / as appended after the extra tags
    assert f[:11] == ["r1", "0", "amp", "1", "255", O.cigar_str(ops), "*", "0", "0", p["read"], "i" * len(read)]
    tags = dict(x.split(":Z:") for x in f[11:])
    rate, m, mm = O.alignment_rate(a["ref_aligned"], a["read_aligned"])
    assert tags == {"ar": "r1", "e1": "ACGT", "rc": "1", "rm": h_f64(H, rate), "rs": h_f64(H, a["score"]), "as": h_f64(H, a["score"])}
    assert [x[:2] for x in f[11:]] == ["ar", "e1", "rc", "rm", "rs", "as"]
    assert tags["as"] == "4"


def test_merge_reads_by_concatenation(H):
    # merge_reads_by_concatenation + orient_sequence, merger.rs:40-126
    def merge(r1, r2, layout):
        buf = C.create_string_buffer(1024)
        n = H.clqh_merge_reads_by_concatenation(r1, len(r1), r2, len(r2) if r2 is not None else 0, layout.encode(), buf, 1024)
        return None if n == 2 ** 64 - 1 else buf.raw[:n]

    assert merge(b"AACC", b"GGTA", "1F,2F") == b"AACCGGTA"                     # ConcatenateBothForward
    assert merge(b"AACC", b"GGTA", "1F,2C") == b"AACC" + O.reverse_complement(b"GGTA")
    assert merge(b"AACC", b"GGTA", "1F,S:NNNN,2R") == b"AACCNNNNATGG"
    assert merge(b"aacc", b"GGTA", "2F,1C") == b"GGTAGGTT"                     # reverse_complement upper-cases
    assert merge(b"AACC", None, "1F") == b"AACC"
    assert merge(b"AACC", None, "1F,2F") is None                               # assert!(reads.read_two.is_some()) panics
    assert merge(b"AACC", b"GG", "1U") is None                                 # Unknown orientation panics


def test_orient_sequence_reference_vectors(H):
    # merger.rs:689-730, through the single-read layout items 1F / 1R / 1C / 1U
    def orient(seq, item):
        buf = C.create_string_buffer(64)
        n = H.clqh_merge_reads_by_concatenation(seq, len(seq), None, 0, item.encode(), buf, 64)
        return None if n == 2 ** 64 - 1 else buf.raw[:n]

    assert orient(b"ACGT", "1F") == b"ACGT"        # :690-695
    assert orient(b"ACGT", "1R") == b"TGCA"        # :697-702
    assert orient(b"ACGT", "1C") == b"ACGT"        # :704-709 (its own reverse complement)
    assert orient(b"AAAA", "1C") == b"TTTT"        # :711-716
    assert orient(b"ACGT", "1U") is None           # :718-723 panics with "Unknown"
    for item in ("1F", "1R", "1C"):                # :725-730
        assert orient(b"", item) == b""


# extend_hit's assertions: linked_alignment.rs:370-412 and :544-582  (search, search_location, reference, reference_location, length)
EXTEND_HIT = [
    (b"ACGTACGT", 0, b"ACGTACGT", 0, 8), (b"ACGTTTTT", 0, b"ACGTACGT", 0, 4), (b"TTTT", 0, b"ACGT", 0, 0),
    (b"TTACGT", 2, b"ACGT", 0, 4), (b"ACGT", 0, b"TTACGT", 2, 4), (b"RCGT", 0, b"ACGT", 0, 0),
    (b"AATGATACGG", 0, b"AATGATACGG", 0, 10), (b"AATGATACGG", 0, b"AATGATACGGAAA", 0, 10),
    (b"AATGATACGG", 0, b"GGAATGATACGGAAA", 2, 10), (b"AATGATACGG", 0, b"AAA", 0, 2),
]


def test_extend_hit_reference_vectors(H):
    orc = O.lib()
    orc.orc_extend_hit.restype = C.c_size_t
    orc.orc_extend_hit.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t]
    for s, sl, r, rl, want in EXTEND_HIT:
        assert H.clqh_extend_hit(s, len(s), sl, r, len(r), rl) == want, (s, r)
        assert orc.orc_extend_hit(s, len(s), sl, r, len(r), rl) == want, (s, r)


def h_consensus(H, a1, q1, a2, q2):
    n = len(a1)
    ob, oq = C.create_string_buffer(max(1, n)), C.create_string_buffer(max(1, n))
    r = H.clqh_alignment_rate_and_consensus(a1, q1, len(q1), a2, q2, len(q2), n, ob, oq)
    return None if r == 2 ** 64 - 1 else (ob.raw[:n], oq.raw[:n])


def test_phred_and_consensus_cpp(H, goldens):
    # utils/read_utils.rs:26-38, merger.rs:428-498 against the reference's vectors and the oracle
    for t in goldens["phred"]["combine"]:
        assert H.clqh_combine_phred_scores(ord(t["a"]), ord(t["b"]), 1 if t["agree"] else 0) == ord(t["out"])
    L = O.lib()
    for a in range(33, 127, 3):
        for b in range(33, 127, 5):
            for agree in (0, 1):
                assert H.clqh_combine_phred_scores(a, b, agree) == L.orc_combine_phred_scores(a, b, agree)
    for m in goldens["mergers"]:
        r = O.align_pair(m["read1"].encode(), m["read2_revcomp"].encode(), m["scoring"], "maxlen")
        got = h_consensus(H, r["ref_aligned"], m["qual1"].encode(), r["read_aligned"], m["qual2_rev"].encode())
        assert got[0].decode() == m["expect_merged"], m["name"]
        assert got == O.alignment_rate_and_consensus(r["ref_aligned"], m["qual1"].encode(), r["read_aligned"], m["qual2_rev"].encode())
    assert h_consensus(H, b"A--", b"H", b"A--", b"H") is None


def test_orientation_cpp(H, goldens):
    # orient_by_longest_segment / find_greedy_non_overlapping_segments (linked_alignment.rs:24-32, :97-130): the reference's
    # two asserted cases, its print-only inputs, and random reads against the oracle (suffix-array order of the seed hits)
    def segs(read, ref, k):
        out = np.zeros((len(read) + 2, 3), np.uint32)
        sp = C.c_size_t()
        n = H.clqh_find_greedy_non_overlapping_segments(read, len(read), ref, len(ref), k, out.ctypes.data, len(out), C.byref(sp))
        return [tuple(int(v) for v in out[i]) for i in range(n)], sp.value

    for t in goldens["orient"]["kats"]:
        s, _ = segs(t["read"].encode(), t["ref"].encode(), t["seed_size"])
        assert len(s) == t["n_segments"] and [x[0] for x in s] == t["search_starts"]
    cases = [(t["read"].encode(), t["ref"].encode(), t["seed_size"]) for t in goldens["orient"]["kats"] + goldens["orient"]["inputs"]]
    rng = np.random.default_rng(17)
    for _ in range(200):
        ref = bytes(rng.choice(list(b"ACGTacgtN"), size=int(rng.integers(0, 200)), p=[.2, .2, .2, .2, .04, .04, .04, .04, .04]).astype(np.uint8))
        ref = ref + ref[:int(rng.integers(0, 40))]          # repeats: several seed hits per k-mer
        read = bytearray(ref[int(rng.integers(0, 30)):]) if rng.random() < 0.8 else bytearray(rng.choice(list(b"ACGT"), size=60).astype(np.uint8))
        for k in (rng.integers(0, len(read), size=3) if len(read) else []):
            read[int(k)] = int(rng.choice(list(b"ACGTN")))
        read = bytes(read) if rng.random() < 0.5 else O.reverse_complement(bytes(read))
        cases.append((read, ref, int(rng.choice([3, 5, 8, 20]))))
    for read, ref, k in cases:
        assert segs(read, ref, k) == O.find_greedy_non_overlapping_segments(read, ref, k), (read, ref, k)
        assert bool(H.clqh_orient_by_longest_segment(read, len(read), ref, len(ref), k)) == O.orient_by_longest_segment(read, ref, k)[0]


def parse_bam(blob):
    """Minimal BAM reader for the tests: BGZF framing checks + header + records -> (header_text, refs, SAM-like field lists)."""
    import gzip
    import struct
    # every BGZF member carries the BC extra field with its own size; the file ends with the 28-byte EOF marker
    off, n_blocks = 0, 0
    while off < len(blob):
        assert blob[off:off + 4] == b"\x1f\x8b\x08\x04" and blob[off + 12:off + 14] == b"BC"
        bsize = struct.unpack_from("<H", blob, off + 16)[0] + 1
        off += bsize; n_blocks += 1
    assert off == len(blob) and blob[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    raw = gzip.decompress(blob)
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, p)[0]; p += 4
        nm = raw[p:p + ln - 1].decode(); p += ln
        refs.append((nm, struct.unpack_from("<i", raw, p)[0])); p += 4
    recs = []
    while p < len(raw):
        bs = struct.unpack_from("<i", raw, p)[0]; p += 4
        e = p + bs
        ref_id, pos, l_name, mapq, bin_, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", raw, p); p += 32
        name = raw[p:p + l_name - 1].decode(); p += l_name
        cig = struct.unpack_from("<%dI" % n_cig, raw, p); p += 4 * n_cig
        seq = "".join("=ACMGRSVTWYHKDBN"[(raw[p + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)); p += (l_seq + 1) // 2
        qual = raw[p:p + l_seq]; p += l_seq
        tags = []
        while p < e:
            t = raw[p:p + 2].decode(); assert raw[p + 2:p + 3] == b"Z"; p += 3
            z = raw.index(b"\0", p)
            tags.append(t + ":Z:" + raw[p:z].decode()); p = z + 1
        assert p == e and (nref, npos, tlen) == (-1, -1, 0)
        cigar = "".join("%d%s" % (c >> 4, "MIDNSHP=X"[c & 15]) for c in cig) or "*"
        fields = [name, str(flag), refs[ref_id][0], str(pos + 1), str(mapq), cigar, "*", "0", "0", seq or "*",
                  "".join(chr(q + 33) for q in qual) or "*"] + tags
        recs.append((fields, bin_))
    return text, refs, recs, n_blocks


def test_bam_wire_format(H, goldens):
    # BamFileAlignmentWriter (alignment_manager.rs:64-209) + to_sam_record (alignment/alignment_matrix.rs:741-771): header text,
    # reference dictionary, binary record, BGZF framing -- the record equals the SAM text line field by field
    p = goldens["pairs"][3]  # affine_alignment_test_favor_non_special_characters: 66M28D26M6D
    ref, read = p["ref"].encode(), p["read"].encode()
    a = O.align_pair(ref, read, p["scoring"], "maxlen")
    ops = np.ascontiguousarray(a["cigar"], dtype=np.uint32)
    buf = C.create_string_buffer(1 << 16)
    n = H.clqh_bam_file(b"amp", b"read7", ref, len(ref), read, len(read), ops.ctypes.data, len(ops), a["score"], b"e1=ACGT;rc=1;ar=read7", buf, 1 << 16)
    assert n > 0
    text, refs, recs, n_blocks = parse_bam(buf.raw[:n])
    assert text == "@HD\tVN:1.6\n@SQ\tSN:amp\tLN:%d\n@CO\tClique processed\n" % len(ref)
    assert refs == [("amp", len(ref))] and n_blocks == 3 and len(recs) == 1
    line = C.create_string_buffer(8192)
    m = H.clqh_sam_line(b"amp", b"read7", ref, len(ref), read, len(read), ops.ctypes.data, len(ops), a["score"], 0, b"e1=ACGT;rc=1;ar=read7", line, 8192)
    assert recs[0][0] == line.raw[:m].decode().split("\t")
    assert recs[0][1] == 4681  # reg2bin(0, 126): the 16 kb bin of the first window
    assert recs[0][0][5] == "66M28D26M6D" and recs[0][0][10] == "i" * len(read)


# ------------------------------------------------------------------------------------------------ GPU: the batch loop in C++
def _write_inputs(tmp, refs, names, reads, fastq=True):
    fa = os.path.join(tmp, "refs.fa")
    with open(fa, "wb") as f:
        for n, r in zip(names, refs):
            f.write(b">" + n + b"\n" + r + b"\n")
    rp = os.path.join(tmp, "reads.fastq" if fastq else "reads.txt")
    with open(rp, "wb") as f:
        for i, r in enumerate(reads):
            if fastq:
                f.write(b"@q%d some comment\n" % i + r + b"\n+\n" + b"F" * len(r) + b"\n")
            else:
                f.write(r + b"\n")
    return fa, rp


def _run_clq_align(tmp, fa, rp, extra=()):
    out = os.path.join(tmp, "out.sam")
    st = os.path.join(tmp, "stats.json")
    cmd = [CLQ_ALIGN, "--refs", fa, "--reads", rp, "--out", out, "--stats-json", st, "--batch", "257", "--cigar-ops-per-read", "256"] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln.rstrip("\n").split("\t") for ln in open(out) if not ln.startswith("@")]
    head = [ln.rstrip("\n") for ln in open(out) if ln.startswith("@")]
    return head, lines, json.load(open(st))


def _expect_lines(H, want, refs, names, reads, qnames, umi="0123456789"):
    """SAM fields the reference's align_reads + to_sam_record would produce, from the oracle's alignment"""
    exp = []
    for i, rd in enumerate(reads):
        if int(want["status"][i]) != 0:
            continue
        ri = int(want["ref_index"][i])
        o, l = int(want["cigar_off"][i]), int(want["cigar_len"][i])
        cig = want["cigar_pool"][o:o + l]
        ra, qa = O.apply_cigar(refs[ri], rd, cig)
        tags = {"rc": "1", "ar": qnames[i], "rm": h_f64(H, O.alignment_rate(ra, qa)[0]), "as": h_f64(H, float(want["score"][i]))}
        tags["rs"] = tags["as"]
        for k, v in O.extract_tagged_sequences(qa, ra).items():
            if chr(k) in umi:
                tags["e" + chr(k)] = v.decode()
        exp.append(([qnames[i], "0", names[ri].decode(), "1", "255", O.cigar_str(cig), "*", "0", "0", rd.decode(), "i" * len(rd)], tags))
    return exp


def _check(lines, exp):
    assert len(lines) == len(exp)
    for got, (fields, tags) in zip(lines, exp):
        assert got[:11] == fields, (got[0], got[:6], fields[:6])
        assert dict(x.split(":Z:") for x in got[11:]) == tags, got[0]


@pytest.mark.gpu
def test_clq_align_single_amplicon(H, tmp_path):
    from clique_b200 import synth
    c = synth.config_c2(1500)
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(1500)]
    reads[7] = reads[7][:40]                       # ragged
    reads[11] = reads[11] + reads[12]              # 600 bp >= 2 * 216: dropped with a warning by align_reads
    refs, names = c["refs"], c["ref_names"]
    fa, rp = _write_inputs(str(tmp_path), refs, names, reads)
    head, lines, st = _run_clq_align(str(tmp_path), fa, rp)
    assert head[1] == "@SQ\tSN:lineage_amplicon\tLN:215" and head[-1] == "@CO\tClique processed"
    keep = [i for i, r in enumerate(reads) if len(r) < 2 * 216]
    rb, ro = O.pack_seqs(refs)
    qb, qo = O.pack_seqs([reads[i] for i in keep])
    want = O.align_batch(rb, ro, qb, qo, c["scoring"], search="fixed", fixed_ref=np.zeros(len(keep), np.int32), band_mode="readlen", threads=8)
    exp = _expect_lines(H, want, refs, names, [reads[i] for i in keep], ["q%d" % i for i in keep])
    _check(lines, exp)
    assert st["reads"] == 1500 and st["dropped"] == 1500 - len(exp) and st["batches"] == 6
    # the fast formatter (BatchView::append_sam_line) and the owned-object path (alignment -> to_sam_record -> to_sam_line)
    # produce the same text, whatever the thread count
    _, slow, _ = _run_clq_align(str(tmp_path), fa, rp, ["--slow-sam"])
    _, threaded, _ = _run_clq_align(str(tmp_path), fa, rp, ["--threads", "5"])
    assert slow == lines and threaded == lines
    # the lineage amplicon has three tag runs (16 + 12 + 12 columns): e0 / e1 / e2 present on every record
    assert all({"e0", "e1", "e2"} <= {x[:2] for x in ln[11:]} for ln in lines)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["quick", "exhaustive"])
def test_clq_align_panel(H, tmp_path, mode):
    from clique_b200 import synth
    c = synth.config_c4(600, search=mode)
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(600)]
    refs, names = c["refs"][:16], c["ref_names"][:16]
    fa, rp = _write_inputs(str(tmp_path), refs, names, reads, fastq=False)
    head, lines, st = _run_clq_align(str(tmp_path), fa, rp, ["--exhaustive"] if mode == "exhaustive" else [])
    rb, ro = O.pack_seqs(refs)
    qb, qo = O.pack_seqs(reads)
    want = O.align_batch(rb, ro, qb, qo, c["scoring"], search=mode, band_mode="readlen", kmer=(8, 4), threads=8)
    exp = _expect_lines(H, want, refs, names, reads, ["read%d" % i for i in range(600)])
    _check(lines, exp)
    assert st["reads"] == 600


@pytest.mark.gpu
def test_clq_align_rust_bio_branch(H, tmp_path):
    """--rust-bio: the reference's current single-reference branch (alignment_functions.rs:544-603) through the C++ batch loop:
    rust-bio CIGARs (PARITY UNPINNED restatement), score written as 0, tags extracted from those alignments."""
    from clique_b200 import synth
    c = synth.config_c2(500)
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(500)]
    refs, names = c["refs"], c["ref_names"]
    fa, rp = _write_inputs(str(tmp_path), refs, names, reads)
    head, lines, st = _run_clq_align(str(tmp_path), fa, rp, ["--rust-bio"])
    assert len(lines) == 500 and st["aligned"] == 500
    for i, (got, rd) in enumerate(zip(lines, reads)):
        want = O.rustbio_global(refs[0], rd)
        ra, qa = O.apply_cigar(refs[0], rd, want["cigar"])
        tags = {"rc": "1", "ar": "q%d" % i, "rm": h_f64(H, O.alignment_rate(ra, qa)[0]), "as": "0", "rs": "0"}
        for k, v in O.extract_tagged_sequences(qa, ra).items():
            if 48 <= k <= 57:
                tags["e" + chr(k)] = v.decode()
        assert got[:11] == ["q%d" % i, "0", "lineage_amplicon", "1", "255", O.cigar_str(want["cigar"]), "*", "0", "0", rd.decode(), "i" * len(rd)], i
        assert dict(x.split(":Z:") for x in got[11:]) == tags, i


@pytest.mark.gpu
def test_clq_align_two_gpus(H, tmp_path):
    """ShardedAligner over two devices (one host thread + context each, shared source, no collective): same records, in input
    order, as the single-GPU run.  Skipped on a one-GPU box."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "clique_b200", "libclq.so"))
    if lib.clq_device_count() < 2:
        pytest.skip("needs two GPUs")
    from clique_b200 import synth
    c = synth.config_c4(2000, search="quick")
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(2000)]
    refs, names = c["refs"][:16], c["ref_names"][:16]
    fa, rp = _write_inputs(str(tmp_path), refs, names, reads)
    _, one, st1 = _run_clq_align(str(tmp_path), fa, rp)
    _, two, st2 = _run_clq_align(str(tmp_path), fa, rp, ["--gpus", "0,1"])
    assert st2["gpus"] == 2 and st1["reads"] == st2["reads"] == 2000
    assert one == two


def _span_case(n=6000):
    """length-sorted mixed-length input (the hard case for a contiguous byte split): C5 reads, shortest first"""
    from clique_b200 import synth
    c = synth.config_c5(n, unique_per_amplicon=32)
    off = c["read_off"]
    lens = (off[1:] - off[:-1]).astype(np.int64)
    order = np.argsort(lens, kind="stable")
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in order]
    qb, qo = O.pack_seqs(reads)
    return c, qb, qo, c["fixed_ref"][order].astype(np.int32)


def _check_span(br, c, qb, qo, fixed, n_check):
    rb, ro = O.pack_seqs(c["refs"])
    idx = np.unique(np.linspace(0, len(qo) - 2, n_check).astype(np.int64))
    reads = [bytes(qb[int(qo[i]):int(qo[i + 1])]) for i in idx]
    sb, so = O.pack_seqs(reads)
    want = O.align_batch(rb, ro, sb, so, c["scoring"], search="fixed", fixed_ref=fixed[idx], band_mode="readlen", threads=8)
    for k, i in enumerate(idx):
        assert int(br.status[i]) == int(want["status"][k]) == 0, (i, int(br.status[i]))
        assert int(br.ref_index[i]) == int(want["ref_index"][k]), i
        assert int(br.score_scaled[i]) == want["score"][k] * br.scale, i
        o, l = int(want["cigar_off"][k]), int(want["cigar_len"][k])
        assert O.cigar_str(br.cigar(int(i))) == O.cigar_str(want["cigar_pool"][o:o + l]), i


@pytest.mark.gpu
def test_align_reads_span_one_gpu():
    """ShardedAligner::align_reads_span on one device: reads in plain host memory, dynamic batches (much smaller than the
    input, so the guided claims, the buffer pool and the CIGAR-pool rebasing all run), records in input order == oracle."""
    from clique_b200 import AffineScoring
    from clique_b200.host import align_reads_span
    c, qb, qo, fixed = _span_case(3000)
    br, st = align_reads_span([0], c["refs"], qb, qo, AffineScoring(*c["scoring"]), fixed_ref=fixed, batch_reads=512, batch_bytes=1 << 20,
                              max_read_len=1 << 15, cigar_ops_per_read=700, fillers_per_device=2, passes=2)
    assert st["reads"] == 3000 and st["aligned"] == 3000 and st["batches"] > 4
    assert int(br.cigar_len.sum()) == len(br.cigar_pool)
    _check_span(br, c, qb, qo, fixed, 40)


@pytest.mark.gpu
def test_align_reads_span_all_gpus():
    """the same stream sharded over every visible GPU by the product dispatcher (one process, one cursor, no collective):
    every device takes part, the union of the records == the single-device run.  Skipped on a one-GPU box."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "clique_b200", "libclq.so"))
    ng = lib.clq_device_count()
    if ng < 2:
        pytest.skip("needs two GPUs")
    from clique_b200 import AffineScoring
    from clique_b200.host import align_reads_span
    c, qb, qo, fixed = _span_case(6000)
    kw = dict(fixed_ref=fixed, batch_reads=256, batch_bytes=1 << 19, max_read_len=1 << 15, cigar_ops_per_read=700, fillers_per_device=2)
    one, _ = align_reads_span([0], c["refs"], qb, qo, AffineScoring(*c["scoring"]), **kw)
    many, st = align_reads_span(list(range(ng)), c["refs"], qb, qo, AffineScoring(*c["scoring"]), **kw)
    assert st["reads"] == 6000 and all(r > 0 for r in st["device_reads"]), st
    assert (one.score_scaled == many.score_scaled).all() and (one.status == many.status).all() and (one.cigar_len == many.cigar_len).all()
    for i in range(0, 6000, 97):
        assert one.cigar_string(i) == many.cigar_string(i), i


@pytest.mark.gpu
def test_merge_read_pairs_by_alignment_gpu(H, goldens):
    """merge_reads_by_alignment (MergeStrategy::Align, merger.rs:348-396) for a batch of read pairs: every pair is its own
    (reference = read1, read = revcomp(read2)) task on the GPU (merger scoring, x0.25 final-gap multiplier), consensus on the
    host.  Against the reference's three merger vectors and the oracle on random overlapping pairs."""
    rng = np.random.default_rng(31)
    sc = goldens["scorings"]["merger"]
    pairs = [(m["read1"].encode(), m["qual1"].encode(), O.reverse_complement(m["read2_revcomp"].encode()), m["qual2_rev"].encode()[::-1])
             for m in goldens["mergers"]]
    for _ in range(300):  # a fragment of 120-280 bp read from both ends with 150 bp reads and a few errors
        frag = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(120, 280))).astype(np.uint8))
        r1 = bytearray(frag[:150]); r2 = bytearray(O.reverse_complement(frag)[:150])
        for r in (r1, r2):
            for k in rng.integers(0, len(r), size=int(rng.integers(0, 4))):
                r[int(k)] = int(rng.choice(list(b"ACGT")))
        q1 = bytes(rng.integers(40, 75, size=len(r1)).astype(np.uint8)); q2 = bytes(rng.integers(40, 75, size=len(r2)).astype(np.uint8))
        pairs.append((bytes(r1), q1, bytes(r2), q2))
    n = len(pairs)
    r1b, o1 = O.pack_seqs([p[0] for p in pairs]); q1b, _ = O.pack_seqs([p[1] for p in pairs])
    r2b, o2 = O.pack_seqs([p[2] for p in pairs]); q2b, _ = O.pack_seqs([p[3] for p in pairs])
    cap = int(o1[-1] + o2[-1]) + 64
    ob, oq = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    rc = H.clqh_merge_read_pairs_by_alignment(0, n, r1b.ctypes.data, q1b.ctypes.data, o1.ctypes.data, r2b.ctypes.data, q2b.ctypes.data,
                                              o2.ctypes.data, sc["match_score"], sc["mismatch_score"], sc["special_character_score"],
                                              sc["gap_open"], sc["gap_extend"], sc["final_gap_multiplier"], ob.ctypes.data, oq.ctypes.data,
                                              cap, off.ctypes.data)
    assert rc == 0, rc
    for i, (r1, q1, r2, q2) in enumerate(pairs):
        want = O.merge_reads_by_alignment(r1, q1, r2, q2, sc)
        got = (bytes(ob[int(off[i]):int(off[i + 1])]), bytes(oq[int(off[i]):int(off[i + 1])]))
        if want is None:
            assert got == (b"", b""), i
        else:
            assert got == want, i
    for i, m in enumerate(goldens["mergers"]):
        assert bytes(ob[int(off[i]):int(off[i + 1])]).decode() == m["expect_merged"], m["name"]


def test_no_cpu_fallback_in_cpp_layer(tmp_path):
    """Without a CUDA device the C++ host layer refuses to run: there is no CPU path behind it (skipped where a GPU exists)."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "clique_b200", "libclq.so"))
    lib.clq_device_count.restype = ctypes.c_int32
    if lib.clq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    fa, rp = _write_inputs(str(tmp_path), [b"ACGTACGTACGT"], [b"r"], [b"ACGTACGT"])
    r = subprocess.run([CLQ_ALIGN, "--refs", fa, "--reads", rp, "--out", os.path.join(str(tmp_path), "o.sam")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr, r.stderr


@pytest.mark.gpu
def test_clq_align_unknown_strand(H, tmp_path):
    """known_strand = false with one reference (alignment_functions.rs:549-558): reads arriving as their reverse complement are
    oriented by orient_by_longest_segment and reverse-complemented before the alignment."""
    from clique_b200 import synth
    c = synth.config_c2(300)
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(300)]
    flipped = [O.reverse_complement(r) if i % 2 else r for i, r in enumerate(reads)]
    refs, names = c["refs"], c["ref_names"]
    # what the reference would align: the read as it is when the forward strand shares strictly more bases, else its reverse
    # complement.  (The greedy seed chain is kept as the reference has it: an accidental 8-mer hit far down the reference blocks
    # every earlier hit, so a few reads are oriented the wrong way -- by the reference too.)
    decided = [O.orient_by_longest_segment(r, refs[0], 8)[0] for r in flipped]
    expect_in = [r if fw else O.reverse_complement(r) for r, fw in zip(flipped, decided)]
    assert sum(1 for i, fw in enumerate(decided) if fw == (i % 2 == 0)) >= 285
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(); d2.mkdir()
    fa, rp = _write_inputs(str(d1), refs, names, expect_in)
    _, fwd, _ = _run_clq_align(str(d1), fa, rp)
    fa2, rp2 = _write_inputs(str(d2), refs, names, flipped)
    _, ori, _ = _run_clq_align(str(d2), fa2, rp2, ["--unknown-strand"])
    assert len(ori) == len(fwd) == 300
    assert ori == fwd


@pytest.mark.gpu
def test_clq_align_bam_output(H, tmp_path):
    """--out x.bam: the same records as the SAM text, in BAM (fast raw-record encoder and the object path), BGZF blocks
    compressed on several threads."""
    from clique_b200 import synth
    c = synth.config_c4(700, search="quick")
    off = c["read_off"]
    reads = [bytes(c["read_bytes"][int(off[i]):int(off[i + 1])]) for i in range(700)]
    refs = [r[:100] + b"0123" + r[104:] for r in c["refs"][:8]]
    names = c["ref_names"][:8]
    fa, rp = _write_inputs(str(tmp_path), refs, names, reads)
    head, sam, _ = _run_clq_align(str(tmp_path), fa, rp)

    def run_bam(extra):
        out = os.path.join(str(tmp_path), "o.bam")
        r = subprocess.run([CLQ_ALIGN, "--refs", fa, "--reads", rp, "--out", out, "--batch", "257", "--cigar-ops-per-read", "256"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        return parse_bam(open(out, "rb").read())

    for extra in ([], ["--threads", "4"], ["--slow-sam"]):
        text, brefs, recs, n_blocks = run_bam(extra)
        assert text.rstrip("\n").split("\n") == head
        assert brefs == [(n.decode(), len(r)) for n, r in zip(names, refs)]
        assert [r[0] for r in recs] == sam, extra


# ------------------------------------------------------------------------------------------------ the span dispatcher's batch cutter (CPU)
def span_claims(H, lens, n_devices, max_reads, max_bytes, order=0, claimers=2):
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    H.clqh_span_claims.restype = C.c_uint64
    H.clqh_span_claims.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p]
    cap = len(lens) + 16
    out = np.zeros(4 * cap, np.uint64)
    res = C.c_int32(-1)
    k = int(H.clqh_span_claims(off.ctypes.data, len(lens), n_devices, claimers, max_reads, max_bytes, order, out.ctypes.data, cap, C.byref(res)))
    assert k <= cap
    return [tuple(int(v) for v in out[4 * i:4 * i + 4]) for i in range(k)], res.value, off


def check_partition(claims, off, n, max_reads, max_bytes):
    """every read in exactly one claim, every claim within the batch capacity (a single oversize read may stand alone)"""
    seen = np.zeros(n, np.int32)
    for lo, hi, lo2, hi2 in claims:
        assert lo < hi and lo2 <= hi2
        seen[lo:hi] += 1
        seen[lo2:hi2] += 1
        reads = (hi - lo) + (hi2 - lo2)
        nbytes = int(off[hi] - off[lo]) + int(off[hi2] - off[lo2])
        assert reads <= max_reads
        assert nbytes <= max_bytes or reads == 1
    assert (seen == 1).all()


def test_span_claims_uniform_stream_goes_front_to_back(H):
    lens = [300] * 100_000
    claims, order, off = span_claims(H, lens, 2, 1 << 16, 1 << 26)
    assert order == 1
    check_partition(claims, off, len(lens), 1 << 16, 1 << 26)
    assert all(lo2 == hi2 for _, _, lo2, hi2 in claims)
    assert [c[0] for c in claims] == sorted(c[0] for c in claims)       # contiguous, in input order
    assert claims[-1][1] == len(lens)
    sizes = [hi - lo for lo, hi, _, _ in claims]
    assert max(sizes[len(sizes) // 2:]) <= max(sizes[:len(sizes) // 2])  # guided: the claims shrink towards the end


@pytest.mark.parametrize("ascending", [True, False])
@pytest.mark.parametrize("n_devices", [1, 2, 8])
def test_span_claims_sorted_stream_mixes_both_ends(H, ascending, n_devices):
    rng = np.random.default_rng(5 + n_devices)
    lens = np.sort(rng.choice([300, 450, 700, 1000, 1500, 2200, 3300, 5000], size=60_000))
    if not ascending:
        lens = lens[::-1]
    max_reads, max_bytes = 1 << 15, 1 << 26
    claims, order, off = span_claims(H, lens, n_devices, max_reads, max_bytes)
    assert order == 3                                                     # auto -> two-ended
    check_partition(claims, off, len(lens), max_reads, max_bytes)
    two = [c for c in claims if c[3] > c[2]]
    assert len(two) >= len(claims) // 2
    lo, hi, lo2, hi2 = claims[0]
    long_mean = float(np.mean(lens[lo:hi])), float(np.mean(lens[lo2:hi2]))
    assert long_mean[0] >= 4000 and long_mean[1] <= 1000                  # first claim: longest reads + shortest reads
    # cost (cells ~ len^2) of the claims decreases overall: the cheap reads are what is left for the end
    cost = [float((lens[a:b].astype(np.float64) ** 2).sum() + (lens[c:d].astype(np.float64) ** 2).sum()) for a, b, c, d in claims]
    assert cost[0] >= cost[-1] and sum(cost[:len(cost) // 2]) > sum(cost[len(cost) // 2:])


def test_span_claims_explicit_orders_and_oversize_read(H):
    lens = np.sort(np.random.default_rng(9).integers(100, 4000, size=20_000))
    for order in (1, 2, 3):
        claims, res, off = span_claims(H, lens, 4, 4096, 1 << 22, order=order)
        assert res == order
        check_partition(claims, off, len(lens), 4096, 1 << 22)
    longest_first, _, _ = span_claims(H, lens, 4, 4096, 1 << 22, order=2)
    assert longest_first[0][1] == len(lens) and all(c[2] == c[3] for c in longest_first)
    # one read larger than a whole batch is handed over alone (the dispatcher reports it CLQ_READ_TOO_LONG)
    lens2 = [200] * 5000 + [1 << 20] + [200] * 5000
    for order in (0, 1, 2, 3):
        claims, _, off = span_claims(H, lens2, 2, 4096, 1 << 18, order=order)
        check_partition(claims, off, len(lens2), 4096, 1 << 18)
        assert any(hi - lo == 1 and lo == 5000 and lo2 == hi2 for lo, hi, lo2, hi2 in claims)
    assert span_claims(H, [], 2, 4096, 1 << 18)[0] == []
