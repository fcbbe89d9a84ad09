#!/usr/bin/env python3
"""Generate tests/golden/reference_goldens.json from the reference's own unit tests and fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is Rust and cannot be executed here, so nothing is *run*: the script lifts the
input sequences and the asserted expectations (string literals) straight out of the reference's
`#[test]` functions by file:line range, and copies the small FASTA fixtures those tests load.
Every vector records the file:line it was taken from.  No reference source code is copied.
"""
import json
import os
import re
import sys

REF = "/root/reference/rust_cmd"
SRC = REF + "/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

LIT = re.compile(r'"((?:[^"\\]|\\.)*)"')


def lines(path, lo, hi):
    with open(os.path.join(SRC, path)) as f:
        all_lines = f.read().split("\n")
    return all_lines[lo - 1:hi]


def lits(path, lo, hi, skip_comments=True):
    """All string literals in path:lo-hi (1-based, inclusive), in order."""
    out = []
    for ln in lines(path, lo, hi):
        s = ln.strip()
        if skip_comments and s.startswith("//"):
            continue
        out.extend(_scan(ln, skip_comments))
    return out


def _scan(ln, skip_comments):
    """String literals of one line; a // outside a literal starts a comment."""
    res, i, n = [], 0, len(ln)
    while i < n:
        c = ln[i]
        if c == '"':
            j = i + 1
            buf = []
            while j < n and ln[j] != '"':
                if ln[j] == "\\" and j + 1 < n:
                    buf.append(ln[j:j + 2]); j += 2
                else:
                    buf.append(ln[j]); j += 1
            res.append("".join(buf))
            i = j + 1
        elif skip_comments and ln.startswith("//", i):
            break
        elif c == "'" and i + 2 < n and ln[i + 2] == "'":
            i += 3  # char literal such as b'"'
        else:
            i += 1
    return res


def floats_after(path, lo, hi, key):
    for ln in lines(path, lo, hi):
        m = re.search(key + r"\s*:\s*(-?[0-9.]+)", ln)
        if m:
            return float(m.group(1))
    raise KeyError(key)


def scoring_at(path, lo, hi):
    keys = ["match_score", "mismatch_score", "special_character_score", "gap_open", "gap_extend", "final_gap_multiplier"]
    return {k: floats_after(path, lo, hi, k) for k in keys}


def read_fasta(path):
    recs = []
    name, seq = None, []
    with open(path) as f:
        for ln in f:
            ln = ln.rstrip("\n").rstrip("\r")
            if ln.startswith(">"):
                if name is not None:
                    recs.append({"name": name, "seq": "".join(seq)})
                name, seq = ln[1:].split()[0], []
            elif ln.strip():
                seq.append(ln.strip())
    if name is not None:
        recs.append({"name": name, "seq": "".join(seq)})
    return recs


def revcomp(s):
    # reverse_complement, utils/read_utils.rs:50-72 (upper-cases, IUPAC aware)
    m = {"A": "T", "T": "A", "G": "C", "C": "G", "R": "Y", "Y": "R", "S": "S", "W": "W", "K": "M", "M": "K",
         "B": "V", "D": "H", "H": "D", "V": "B", "N": "N"}
    return "".join(m.get(c.upper(), c) for c in reversed(s))


DEFAULT_DNA = scoring_at("alignment/scoring_functions.rs", 77, 86)
MERGER = scoring_at("merger.rs", 130, 139)
AM = "alignment/alignment_matrix.rs"

pairs = []


def pair(name, cite, ref, read, scoring, band, **expect):
    pairs.append({"name": name, "cite": cite, "ref": ref, "read": read, "scoring": scoring, "band_mode": band,
                  "expect": expect})


# --- global affine string goldens, alignment/alignment_matrix.rs ---
L = lits(AM, 1195, 1215)
pair("affine_special_scoring_test", AM + ":1195-1215", L[0], L[1], scoring_at(AM, 1199, 1206), "maxlen",
     ref_aligned=L[4], read_aligned=L[5])
L = lits(AM, 1253, 1273)
pair("affine_special_practical_test", AM + ":1253-1273", L[0], L[1], scoring_at(AM, 1258, 1265), "maxlen",
     ref_aligned=L[4], read_aligned=L[5])
L = lits(AM, 1276, 1296)
pair("affine_alignment_test", AM + ":1276-1296", L[0], L[1], scoring_at(AM, 1281, 1288), "maxlen",
     ref_aligned=L[4], read_aligned=L[5])
L = lits(AM, 1298, 1315)
pair("affine_alignment_test_favor_non_special_characters", AM + ":1298-1315", L[0], L[1], DEFAULT_DNA, "maxlen",
     ref_aligned=L[5], read_aligned=L[6])
L = lits(AM, 1652, 1664)
pair("test_identical_sequences_global_alignment", AM + ":1652-1664", L[0], L[0], DEFAULT_DNA, "maxlen",
     ref_aligned=L[0], read_aligned=L[0], cigar="8M")
L = lits(AM, 1666, 1688)
pair("test_single_base_deletion", AM + ":1666-1688", L[0], L[1], scoring_at(AM, 1670, 1677), "maxlen", total_del=1)
L = lits(AM, 1690, 1713)
pair("test_single_base_insertion", AM + ":1690-1713", L[0], L[1], scoring_at(AM, 1694, 1701), "maxlen", total_ins=1)

# --- soft-clip realignment goldens, extractor.rs:740-784 (align_two_strings with default_dna, extractor.rs:149,162) ---
L = lits("extractor.rs", 740, 759)
read, reference = L[0], L[1]
start_pos, clip = 23, 9  # :743, :745
ref_pos = start_pos - 1
pair("test_recover_align_sequences/leading_softclip", "extractor.rs:740-759", reference[:ref_pos], read[:clip],
     DEFAULT_DNA, "maxlen", read_aligned=L[3][:ref_pos], ref_aligned=L[4][:ref_pos])
assert L[3][:ref_pos].replace("-", "") == read[:clip] and L[4][:ref_pos] == reference[:ref_pos]
L = lits("extractor.rs", 762, 784)
read, reference = L[0], L[1]
# cigar 38M 4I 54M 2S from start_pos 14 (:764-769)
ref_pos = 14 - 1 + 38 + 54
read_pos = 38 + 4 + 54
exp_read, exp_ref = L[3], L[4]
n_tail = max(len(reference) - ref_pos, 2)
assert exp_ref[-n_tail:].replace("-", "") == reference[ref_pos:] and exp_read[-n_tail:].replace("-", "") == read[read_pos:read_pos + 2]
pair("test_recover_align_sequences/trailing_softclip", "extractor.rs:762-784", reference[ref_pos:], read[read_pos:read_pos + 2],
     DEFAULT_DNA, "maxlen", read_aligned=exp_read[-n_tail:], ref_aligned=exp_ref[-n_tail:])

# --- read merging goldens, merger.rs:527-580: align(read1, revcomp(read2)) with the x0.25 scoring, then
#     alignment_rate_and_consensus (merger.rs:428-498) restated in the test ---
mergers = []
for (lo, hi, nm) in [(527, 544, "read_merger_simple"), (547, 563, "read_merger_real_from_palincode"),
                     (566, 580, "read_merger_simple_no_merge")]:
    L = lits("merger.rs", lo, hi)
    r1, q1, r2, q2 = L[0], L[1], L[2], L[3]
    merged = L[-1]
    mergers.append({"name": nm, "cite": "merger.rs:%d-%d" % (lo, hi), "read1": r1, "qual1": q1,
                    "read2_revcomp": revcomp(r2), "qual2_rev": q2[::-1], "scoring": MERGER, "band_mode": "maxlen",
                    "expect_merged": merged})

# --- best-reference goldens, alignment_functions.rs:931-1073 ---
AF = "alignment_functions.rs"
CLI1 = scoring_at(AF, 964, 971)
L1 = lits(AF, 931, 1011)
best_ref = [
    {"name": "test_find_best_reference/read1", "cite": AF + ":931-989", "fasta": "test_best_alignment.fasta",
     "read": L1[1].upper(), "scoring": CLI1, "kmer": [8, 8], "expect_ref_name": [s for s in L1 if s.startswith("1_")][0]},
    {"name": "test_find_best_reference/read2", "cite": AF + ":991-1011", "fasta": "test_best_alignment.fasta",
     "read": [s for s in L1 if s.startswith("atgg")][1].upper(), "scoring": CLI1, "kmer": [8, 8],
     "expect_ref_name": [s for s in L1 if s.startswith("2_")][0]},
]
L2 = lits(AF, 1014, 1073)
best_ref.append({"name": "test_find_best_reference2", "cite": AF + ":1014-1073", "fasta": "test_ref_alignment.fasta",
                 "read": [s for s in L2 if s.startswith("ATGG")][0].upper(), "scoring": scoring_at(AF, 1046, 1053),
                 "kmer": [8, 8], "expect_ref_name": [s for s in L2 if s.startswith("ref_")][0]})

fastas = {}
for fn in ["test_best_alignment.fasta", "test_ref_alignment.fasta", "two_references.fa", "two_references_just_one.fa"]:
    fastas[fn] = read_fasta(REF + "/test_data/" + fn)
# first 64 records of the 180-reference panel (bench config C4)
panel = read_fasta(REF + "/test_data/18guide1_pcr_sequence.fasta")
assert len(panel) == 180  # reference/fasta_reference.rs:226-236
fastas["18guide1_pcr_sequence.first64"] = panel[:64]

# --- three_way_max_and_direction truth table, alignment/alignment_matrix.rs:1544-1592 ---
tie_table = [
    {"up": 10.0, "left": 5.0, "diag": 3.0, "val": 10.0, "dir": "Up"},
    {"up": 3.0, "left": 10.0, "diag": 5.0, "val": 10.0, "dir": "Left"},
    {"up": 3.0, "left": 5.0, "diag": 10.0, "val": 10.0, "dir": "Diag"},
    {"up": 10.0, "left": 5.0, "diag": 10.0, "val": 10.0, "dir": "Diag"},
    {"up": 5.0, "left": 10.0, "diag": 10.0, "val": 10.0, "dir": "Diag"},
    {"up": 7.0, "left": 7.0, "diag": 7.0, "val": 7.0, "dir": "Diag"},
    {"up": -10.0, "left": -5.0, "diag": -3.0, "val": -3.0, "dir": "Diag"},
]
# sanity: the numbers above must literally appear in the cited lines
txt = "\n".join(lines(AM, 1544, 1592))
for t in tie_table:
    pat = r"three_way_max_and_direction\(&%s, &%s, &%s\)" % (t["up"], t["left"], t["diag"])
    assert re.search(pat.replace(".", r"\."), txt), pat

# --- match_mismatch table, alignment/scoring_functions.rs:232-258 (default_dna) ---
mm_table = []
for ln in lines("alignment/scoring_functions.rs", 232, 258):
    m = re.search(r"match_mismatch\(&b'(.)', &b'(.)'\), (-?[0-9.]+)\)", ln)
    if m:
        mm_table.append({"a": m.group(1), "b": m.group(2), "score": float(m.group(3))})
assert len(mm_table) == 10

# --- simplify_cigar_string, alignment_functions.rs:1075-1147 (twin: alignment_manager.rs:429-559) ---
simplify = [
    {"in": "1M1M1M", "out": "3M"}, {"in": "1M1I1M1M", "out": "1M1I2M"}, {"in": "", "out": ""},
    {"in": "5D", "out": "5D"}, {"in": "3M2D1I4M", "out": "3M2D1I4M"}, {"in": "1D2D3D", "out": "6D"},
    {"in": "1I1I1I", "out": "3I"},
]

# --- k-mer index, reference/fasta_reference.rs:226-264 ---
kmers = [
    {"fasta": "two_references_just_one.fa", "k": 15, "skip": 5, "n_refs": 1,
     "contains": [["cas_tag", "GGGCGAGATCAAGCA"]], "not_contains": []},
    {"fasta": "two_references.fa", "k": 15, "skip": 5, "n_refs": 2,
     "contains": [["cas_tag", "TTTTTTTTTTTTTTC"], ["v10", "AAAAAAAAAAAATTC"]],
     "not_contains": [["cas_tag", "TCACCTATTAGCGGCTAA"], ["v10", "TCACCTATTAGCGGCTAA"]]},
]

# --- get_reference_alignment_rate, consensus/consensus_builders.rs:771-795, :1059-1080 (the `rm` tag) ---
CB = "consensus/consensus_builders.rs"
alignment_rate = []
for lo, hi in ((771, 796), (1059, 1080)):
    block_all = "\n".join(lines(CB, lo, hi))
    for block in re.split(r"fn test_", block_all)[1:]:      # one scope per #[test] function
        strs = {}
        for m in re.finditer(r"let (\w+) = b\"([^\"]*)\";", block):
            strs[m.group(1)] = m.group(2)
        for m in re.finditer(r"get_reference_alignment_rate\((\w+), (\w+)\)", block):
            ref_v, read_v = m.group(1), m.group(2)
            tail = block[m.end():m.end() + 200]
            val = re.search(r"(?:assert_eq!\(\w+, |^, )([0-9.]+)\)", tail, flags=re.M) or re.search(r", ([0-9.]+)\)", tail)
            alignment_rate.append({"ref": strs[ref_v], "read": strs[read_v], "rate": float(val.group(1)), "cite": "%s:%d-%d" % (CB, lo, hi)})
assert len(alignment_rate) == 8, alignment_rate


# --- extract_tagged_sequences, extractor.rs:491-546 and :642-667 (the e0..e9 BAM tags of align_reads) ---
EX = "extractor.rs"
tagged = []
L = lits(EX, 492, 507)
tagged.append({"name": "tagged_sequence_test_space", "cite": EX + ":492-507", "ref": L[0], "read": L[1], "expect": {"1": L[2]}})
L = lits(EX, 510, 518)
tagged.append({"name": "test_real_example", "cite": EX + ":510-518", "ref": L[0], "read": L[1], "expect": {"1": L[3]}})
L = lits(EX, 521, 547)
tagged.append({"name": "lower_and_uppercase_test", "cite": EX + ":521-547", "ref": L[0], "read": L[1], "expect": {"A": L[3], "a": L[4]}})
L = lits(EX, 643, 648)
tagged.append({"name": "test_extract_tagged_sequences_basic", "cite": EX + ":643-648", "ref": L[0], "read": L[1], "expect": {"0": L[2]}})
L = lits(EX, 651, 657)
tagged.append({"name": "test_extract_tagged_sequences_multiple_tags", "cite": EX + ":651-657", "ref": L[0], "read": L[1],
               "expect": {"0": L[2], "1": L[3]}})
L = lits(EX, 660, 667)
tagged.append({"name": "test_extract_tagged_sequences_uppercase_tracking", "cite": EX + ":660-667", "ref": L[0], "read": L[1],
               "expect_keys": ["A", "a"]})
# (test_real_example's strings differ in length: the function zips them and stops at the shorter)

# --- reverse_complement, utils/read_utils.rs:141-198 ---
RU = "utils/read_utils.rs"
revcomp_kats = []
for ln in lines(RU, 141, 190):
    m = re.search(r'reverse_complement\(b"([^"]*)"\), b"([^"]*)"\)', ln)
    if m:
        revcomp_kats.append({"in": m.group(1), "out": m.group(2)})
assert len(revcomp_kats) == 24, len(revcomp_kats)

# --- phred helpers, utils/read_utils.rs:119-140 ---
phred = {"to_prob": [], "to_phred": [], "combine": []}
for ln in lines(RU, 119, 140):
    m = re.search(r"phred_to_prob\(&b'(.)'\), ([0-9.e-]+)\)", ln)
    if m:
        phred["to_prob"].append({"phred": m.group(1), "prob": float(m.group(2))})
    m = re.search(r"prob_to_phred\(([0-9.e-]+)\), b'(.)'\)", ln)
    if m:
        phred["to_phred"].append({"prob": float(m.group(1)), "phred": m.group(2)})
    m = re.search(r"combine_phred_scores\(&b'(.)',&b'(.)', (true|false)\), b'(.)'\)", ln)
    if m:
        phred["combine"].append({"a": m.group(1), "b": m.group(2), "agree": m.group(3) == "true", "out": m.group(4)})
assert (len(phred["to_prob"]), len(phred["to_phred"]), len(phred["combine"])) == (4, 4, 2), phred

# --- orient_by_longest_segment / find_greedy_non_overlapping_segments, linked_alignment.rs:520-540 (+ the print-only inputs of
#     :591-625 as extra inputs for the C++-vs-oracle comparison) ---
LA = "linked_alignment.rs"
L = lits(LA, 521, 540)
orient = {"kats": [
    {"cite": LA + ":521-529", "ref": L[0], "read": L[1], "seed_size": 5, "n_segments": 1, "search_starts": [0]},
    {"cite": LA + ":531-540", "ref": L[2], "read": L[3], "seed_size": 5, "n_segments": 2, "search_starts": [0, 18]}],
    "inputs": []}
for lo, hi in ((592, 603), (606, 616)):
    L = lits(LA, lo, hi)
    orient["inputs"].append({"cite": "%s:%d-%d" % (LA, lo, hi), "ref": L[0], "read": L[1], "seed_size": 20})

# --- ConvexScoring::gap KATs, alignment/scoring_functions.rs:200-213 ---
convex_gap = [{"gap_open": -10.0, "len": 1, "gap": -10.0}, {"gap_open": -10.0, "len": 10, "gap": -9.0}]

# --- amplicons used by the synthetic bench configs ---
tc = lines("temp_compare", 12, 12)[0].replace(" ", "")
amplicon_c2 = tc
assert len(amplicon_c2) == 215, len(amplicon_c2)
long_ref = lits(AM, 1376, 1376)[0]
amplicon_c3 = long_ref[:1000]

# --- survey-run KATs (restated oracle, fresh matrix) -- secondary, not reference-asserted ---
survey_kats = {
    "affine_alignment_test": {"score": 4.0, "cigar": "2M1I2M"},
    "affine_alignment_test_favor_non_special_characters": {"score": 391.5, "cigar": "66M28D26M6D"},
    "read_merger_simple": {"score": 13.75, "cigar": "35D5M35I"},
    "test_find_best_reference/read1": {"scores": [979, 833, 800, 792]},
    "test_find_best_reference/read2": {"scores": [760, 1100, 822, 776]},
    "test_find_best_reference2": {"scores": [1265, 1257, 1275, 1265, 1292, 1346]},
}

out = {
    "_generated_by": "tests/golden/make_golden.py from /root/reference/rust_cmd (unit-test literals + test_data fixtures)",
    "scorings": {"default_dna": DEFAULT_DNA, "merger": MERGER,
                 "cli": scoring_at(AF, 104, 111)},
    "pairs": pairs, "mergers": mergers, "best_ref": best_ref, "fastas": fastas, "tie_table": tie_table,
    "match_mismatch_default_dna": mm_table, "simplify_cigar": simplify, "kmers": kmers, "convex_gap": convex_gap, "alignment_rate": alignment_rate,
    "amplicon_c2": amplicon_c2, "amplicon_c3": amplicon_c3, "survey_kats": survey_kats,
    "tagged_sequences": tagged, "reverse_complement": revcomp_kats, "phred": phred, "orient": orient,
}
with open(OUT, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(pairs), "pairs,", len(best_ref), "best-ref,", len(mergers), "mergers")
