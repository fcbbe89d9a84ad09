"""Pin the CPU oracle against every golden vector the reference's own unit tests hold for the hot path
(SURVEY.md section 8c).  Vectors: tests/golden/reference_goldens.json (made by tests/golden/make_golden.py)."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O


def test_tie_break_table(goldens):
    # alignment/alignment_matrix.rs:1544-1592
    names = {0: "Up", 1: "Left", 2: "Diag"}
    for t in goldens["tie_table"]:
        d = C.c_int()
        v = O.lib().orc_three_way_max(t["up"], t["left"], t["diag"], C.byref(d))
        assert v == t["val"] and names[d.value] == t["dir"], t


def test_match_mismatch_table(goldens):
    # alignment/scoring_functions.rs:232-258
    sc = O.affine(goldens["scorings"]["default_dna"])
    for t in goldens["match_mismatch_default_dna"]:
        assert O.lib().orc_match_mismatch(C.byref(sc), ord(t["a"]), ord(t["b"])) == t["score"], t


def test_convex_gap_kats(goldens):
    # alignment/scoring_functions.rs:200-213
    for t in goldens["convex_gap"]:
        assert O.lib().orc_convex_gap(t["gap_open"], t["len"]) == t["gap"]


def test_simplify_cigar(goldens):
    # alignment_functions.rs:1075-1147, alignment_manager.rs:429-559
    for t in goldens["simplify_cigar"]:
        inp = O.cigar_parse(t["in"])
        out = np.zeros(len(inp) + 1, np.uint32)
        n = O.lib().orc_simplify_cigar(inp.ctypes.data if len(inp) else None, len(inp), out.ctypes.data)
        assert O.cigar_str(out[:n]) == t["out"], t


def _check_pair_expect(p, r, ref, read):
    e = p["expect"]
    assert r["status"] == O.OK
    if "ref_aligned" in e:
        assert r["ref_aligned"].decode() == e["ref_aligned"], p["name"]
    if "read_aligned" in e:
        assert r["read_aligned"].decode() == e["read_aligned"], p["name"]
    if "cigar" in e:
        assert O.cigar_str(r["cigar"]) == e["cigar"]
    if "total_del" in e:
        assert sum(int(o) >> 4 for o in r["cigar"] if int(o) & 0xF == 2) == e["total_del"]
    if "total_ins" in e:
        assert sum(int(o) >> 4 for o in r["cigar"] if int(o) & 0xF == 1) == e["total_ins"]
    # the gapped strings are a pure function of CIGAR + sequences (what the host layer rebuilds)
    ra, qa = O.apply_cigar(ref, read, r["cigar"])
    assert ra == r["ref_aligned"] and qa == r["read_aligned"]


def test_pair_goldens_f64(goldens):
    for p in goldens["pairs"]:
        ref, read = p["ref"].encode(), p["read"].encode()
        r = O.align_pair(ref, read, p["scoring"], p["band_mode"])
        _check_pair_expect(p, r, ref, read)


def test_pair_goldens_int_equals_f64(goldens):
    for p in goldens["pairs"]:
        ref, read = p["ref"].encode(), p["read"].encode()
        rc, sci = O.affine_int(p["scoring"])
        assert rc == O.OK
        r = O.align_pair(ref, read, p["scoring"], p["band_mode"])
        ri = O.align_pair_int(ref, read, sci, p["band_mode"])
        assert ri["score_scaled"] == r["score"] * sci.scale
        assert O.cigar_str(ri["cigar"]) == O.cigar_str(r["cigar"])


def test_survey_kats(goldens):
    k = goldens["survey_kats"]
    by = {p["name"]: p for p in goldens["pairs"]}
    for nm in ("affine_alignment_test", "affine_alignment_test_favor_non_special_characters"):
        p = by[nm]
        r = O.align_pair(p["ref"].encode(), p["read"].encode(), p["scoring"], p["band_mode"])
        assert r["score"] == k[nm]["score"] and O.cigar_str(r["cigar"]) == k[nm]["cigar"]


def _merge(a1, q1, a2, q2):
    # alignment_rate_and_consensus, merger.rs:428-498 (bases only)
    out, p1, p2 = bytearray(), 0, 0
    for a, b in zip(a1, a2):
        if a == b:
            out.append(a); p1 += 1; p2 += 1
        elif a == ord("-"):
            out.append(b); p2 += 1
        elif b == ord("-"):
            out.append(a); p1 += 1
        else:
            out.append(a if q1[p1] >= q2[p2] else b); p1 += 1; p2 += 1
    return bytes(out)


def test_merger_goldens(goldens):
    # merger.rs:527-580 -> merge_fasta_bases_by_alignment -> align_two_strings (x0.25 multiplier)
    for m in goldens["mergers"]:
        r = O.align_pair(m["read1"].encode(), m["read2_revcomp"].encode(), m["scoring"], "maxlen")
        assert r["status"] == O.OK
        merged = _merge(r["ref_aligned"], m["qual1"].encode(), r["read_aligned"], m["qual2_rev"].encode())
        assert merged.decode() == m["expect_merged"], m["name"]
    r = O.align_pair(goldens["mergers"][0]["read1"].encode(), goldens["mergers"][0]["read2_revcomp"].encode(),
                     goldens["mergers"][0]["scoring"], "maxlen")
    k = goldens["survey_kats"]["read_merger_simple"]
    assert r["score"] == k["score"] and O.cigar_str(r["cigar"]) == k["cigar"]


def _panel(goldens, name):
    recs = goldens["fastas"][name]
    return [r["name"] for r in recs], O.pack_seqs([r["seq"].encode() for r in recs])


@pytest.mark.parametrize("tb_all", [True, False])
def test_best_reference_goldens(goldens, tb_all):
    # alignment_functions.rs:931-1073: exhaustive_alignment_search, bandwidth = read.len(), last maximum wins
    for t in goldens["best_ref"]:
        names, (rb, ro) = _panel(goldens, t["fasta"])
        qb, qo = O.pack_seqs([t["read"].encode()])
        out = O.align_batch(rb, ro, qb, qo, t["scoring"], search="exhaustive", band_mode="readlen", traceback_all=tb_all)
        assert out["status"][0] == O.OK
        assert names[out["ref_index"][0]] == t["expect_ref_name"], t["name"]
        # per-candidate scores (survey-run KATs)
        want = goldens["survey_kats"][t["name"]]["scores"]
        got = []
        for i in range(len(names)):
            o1 = O.align_batch(rb, ro, qb, qo, t["scoring"], search="fixed", fixed_ref=[i], band_mode="readlen")
            got.append(o1["score"][0])
        assert got == [float(w) for w in want]
        assert out["score"][0] == max(got)


def test_quick_search_agrees_on_goldens(goldens):
    # alignment_functions.rs:693-767 with the test's k=8, skip=8
    for t in goldens["best_ref"]:
        names, (rb, ro) = _panel(goldens, t["fasta"])
        qb, qo = O.pack_seqs([t["read"].encode()])
        out = O.align_batch(rb, ro, qb, qo, t["scoring"], search="quick", band_mode="readlen", kmer=tuple(t["kmer"]))
        assert names[out["ref_index"][0]] == t["expect_ref_name"], t["name"]


def test_kmer_index_goldens(goldens):
    # reference/fasta_reference.rs:226-264
    L = O.lib()
    for t in goldens["kmers"]:
        names, (rb, ro) = _panel(goldens, t["fasta"])
        assert len(names) == t["n_refs"]
        ix = L.orc_kmer_index_build(len(names), rb.ctypes.data, ro.ctypes.data, t["k"], t["skip"])
        try:
            for nm, kmer in t["contains"]:
                votes = np.zeros(len(names), np.uint32)
                tot = L.orc_kmer_votes(ix, kmer.encode(), len(kmer), votes.ctypes.data)
                assert tot == 1 and votes[names.index(nm)] == 1, (nm, kmer)
            for nm, kmer in t["not_contains"]:
                votes = np.zeros(len(names), np.uint32)
                L.orc_kmer_votes(ix, kmer[:t["k"]].encode(), t["k"], votes.ctypes.data)
                assert votes[names.index(nm)] == 0
        finally:
            L.orc_kmer_index_free(ix)


def test_panel_fixture(goldens):
    recs = goldens["fastas"]["18guide1_pcr_sequence.first64"]
    assert len(recs) == 64 and all(len(r["seq"]) == 302 for r in recs)
    assert len(goldens["amplicon_c2"]) == 215 and len(goldens["amplicon_c3"]) == 1000


def test_alignment_rate_goldens(goldens):
    # get_reference_alignment_rate, consensus/consensus_builders.rs:771-795, :1059-1080
    for t in goldens["alignment_rate"]:
        r, m, mm = O.alignment_rate(t["ref"].encode(), t["read"].encode())
        assert r == t["rate"], t
    # the traceback reports the same counters as the function applied to its gapped strings
    for p in goldens["pairs"]:
        a = O.align_pair(p["ref"].encode(), p["read"].encode(), p["scoring"], "maxlen")
        r, m, mm = O.alignment_rate(a["ref_aligned"], a["read_aligned"])
        assert (a["matches"], a["mismatches"]) == (m, mm)


def test_extract_tagged_sequences_goldens(goldens):
    # extractor.rs:491-546, :642-667
    for t in goldens["tagged_sequences"]:
        got = O.extract_tagged_sequences(t["read"].encode(), t["ref"].encode())
        for k, v in t.get("expect", {}).items():
            assert got[ord(k)].decode() == v, (t["name"], k)
        for k in t.get("expect_keys", []):
            assert ord(k) in got, (t["name"], k)


def test_reverse_complement_goldens(goldens):
    # utils/read_utils.rs:141-198
    for t in goldens["reverse_complement"]:
        assert O.reverse_complement(t["in"].encode()).decode() == t["out"], t
    seq = b"ACGTRYSWKMBDHVN"
    assert O.reverse_complement(O.reverse_complement(seq)) == seq


def test_phred_goldens(goldens):
    # utils/read_utils.rs:119-140
    L = O.lib()
    for t in goldens["phred"]["to_prob"]:
        assert L.orc_phred_to_prob(ord(t["phred"])) == t["prob"], t
    for t in goldens["phred"]["to_phred"]:
        assert L.orc_prob_to_phred(t["prob"]) == ord(t["phred"]), t
    for t in goldens["phred"]["combine"]:
        assert L.orc_combine_phred_scores(ord(t["a"]), ord(t["b"]), 1 if t["agree"] else 0) == ord(t["out"]), t
    for ph in b"!+5I":  # test_phred_roundtrip, :259-266
        assert L.orc_prob_to_phred(L.orc_phred_to_prob(ph)) == ph
    assert L.orc_phred_to_prob(ord("!")) == 1.0


def test_merge_reads_by_alignment_goldens(goldens):
    # merger.rs:527-580 through the C restatement of alignment_rate_and_consensus (bases AND qualities)
    for m in goldens["mergers"]:
        r = O.align_pair(m["read1"].encode(), m["read2_revcomp"].encode(), m["scoring"], "maxlen")
        bases, quals = O.alignment_rate_and_consensus(r["ref_aligned"], m["qual1"].encode(), r["read_aligned"], m["qual2_rev"].encode())
        assert bases.decode() == m["expect_merged"], m["name"]
        assert bases == _merge(r["ref_aligned"], m["qual1"].encode(), r["read_aligned"], m["qual2_rev"].encode())
        assert len(quals) == len(bases)
    # quality rules: agreement multiplies the error probabilities, a gap passes the other read's quality through
    b, q = O.alignment_rate_and_consensus(b"AC-T", b"HHH", b"ACG-", b"+++")
    assert b == b"ACGT" and q == bytes([O.lib().orc_combine_phred_scores(72, 43, 1)] * 2) + b"+H"
    assert O.alignment_rate_and_consensus(b"A--", b"H", b"A--", b"H") is None  # gap/gap consumes qualities: the reference panics


def test_orientation_goldens(goldens):
    # linked_alignment.rs:520-540
    for t in goldens["orient"]["kats"]:
        segs, start = O.find_greedy_non_overlapping_segments(t["read"].encode(), t["ref"].encode(), t["seed_size"])
        assert len(segs) == t["n_segments"] and [s[0] for s in segs] == t["search_starts"], t
        assert O.orient_by_longest_segment(t["read"].encode(), t["ref"].encode(), t["seed_size"])[0] is True
    for t in goldens["orient"]["inputs"]:  # print-only tests of the reference: sanity properties
        ref, read = t["ref"].encode(), t["read"].encode()
        segs, start = O.find_greedy_non_overlapping_segments(read, ref, t["seed_size"])
        assert segs and start == min(s[1] for s in segs)
        for ss, rs, ln in segs:
            assert read[ss:ss + ln].upper() == ref[rs:rs + ln].upper() and ln >= t["seed_size"]
        fwd, f, r = O.orient_by_longest_segment(read, ref, t["seed_size"])
        assert fwd and f == sum(s[2] for s in segs) and r < f
        assert O.orient_by_longest_segment(O.reverse_complement(read), ref, t["seed_size"])[0] is False
