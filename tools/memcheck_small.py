"""Small run over every kernel family for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/memcheck_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _oracle as O
from clique_b200 import AffineScoring, Aligner, Reference, ReferenceManager, RustBioScoring, TwoPieceScoring
from clique_b200.aligner import pack_reads

rng = np.random.default_rng(1)
rs = lambda n, a=b"ACGT": bytes(rng.choice(list(a), size=n).astype(np.uint8))


def mut(s):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < 0.03:
            continue
        if r < 0.06:
            out += rs(2)
        out.append(c)
    return bytes(out)


al = Aligner(device=0, max_reads=256, max_read_bytes=1 << 20, max_read_len=4096, cigar_ops_per_read=128, n_slots=2)
cli, dna = AffineScoring.align_reads_default(), AffineScoring.default_dna()
refs = [rs(120), rs(90, b"ACGTN"), rs(150, b"ACGTacgtN#")]
reads = [mut(refs[i % 3]) for i in range(37)] + [b"", rs(5), rs(700)]
bad = 0
for name, refset, sc, search, band, fixed in [
    ("pack/fixed", refs[:1], cli, "fixed", "readlen", np.zeros(len(reads), np.int32)),
    ("fast/fixed-multi", refs[:2], cli, "fixed", "readlen", np.arange(len(reads), dtype=np.int32) % 2),
    ("generic/fin", refs[:1], dna, "fixed", "maxlen", np.zeros(len(reads), np.int32)),
    ("generic/alphabet", refs, cli, "exhaustive", "readlen", None),
    ("pack/exhaustive", refs[:2], cli, "exhaustive", "readlen", None),
    ("quick", refs[:2], cli, "quick", "readlen", None),
]:
    al.set_references(ReferenceManager([Reference(r, b"r%d" % i) for i, r in enumerate(refset)]))
    qb, qo = pack_reads(reads)
    for cfg in (-1, 0, 3, 5):     # automatic geometry, the narrowest one (multi-stripe for the 700 bp read), G = 16 / 32 (narrow last stripe)
        al.set_option("force_cfg", cfg)
        br = al.align_batch(qb, qo, sc, search, band, fixed_ref=fixed, with_stats=True, extract_tags=True)
        so = al.align_batch(qb, qo, sc, search, band, fixed_ref=fixed, score_only=True)
        rb, ro = O.pack_seqs(refset)
        want = O.align_batch(rb, ro, qb, qo, (sc.match_score, sc.mismatch_score, sc.special_character_score, sc.gap_open, sc.gap_extend,
                                               sc.final_gap_multiplier), search=search, fixed_ref=fixed, band_mode=band, threads=4)
        for i in range(len(reads)):
            o, l = int(want["cigar_off"][i]), int(want["cigar_len"][i])
            if int(br.status[i]) != int(want["status"][i]) or int(br.score_scaled[i]) != want["score"][i] * br.scale or \
                    int(so.score_scaled[i]) != int(br.score_scaled[i]) or \
                    (int(want["status"][i]) == 0 and O.cigar_str(br.cigar(i)) != O.cigar_str(want["cigar_pool"][o:o + l])):
                bad += 1
        print(name, "cfg", cfg, "variant", br.stats["variant"] & 15, "launches", br.stats["launches"])
    al.set_option("force_cfg", -1)
# rust-bio single-reference branch (PACK and int32), tags on
for no_pack in (0, 1):
    al.set_option("no_pack", no_pack)
    tagref = rs(40) + b"0000" + rs(30) + b"111" + rs(40, b"ACGTN")
    al.set_references(ReferenceManager([Reference(tagref, b"t")]))
    rr = [mut(tagref.replace(b"0", b"A").replace(b"1", b"C")) for _ in range(33)] + [b"", rs(3), rs(200, b"ACGTN")]
    qb, qo = pack_reads(rr)
    for cfg in (-1, 0, 3, 5):
        al.set_option("force_cfg", cfg)
        br = al.align_batch(qb, qo, RustBioScoring(), "fixed", "maxlen", fixed_ref=np.zeros(len(rr), np.int32), extract_tags=True)
        for i, rd in enumerate(rr):
            w = O.rustbio_global(tagref, rd)
            if w["score"] != int(br.score_scaled[i]) or O.cigar_str(w["cigar"]) != br.cigar_string(i):
                bad += 1
    al.set_option("force_cfg", -1)
al.set_option("no_pack", 0)
print("rustbio ok")
cv = TwoPieceScoring(10, -9, 9, -20, -2, -40, -1)
al.set_references(ReferenceManager([Reference(refs[0], b"r0")]))
qb, qo = pack_reads(reads)
br = al.align_batch(qb, qo, cv, "fixed", "readlen", fixed_ref=np.zeros(len(reads), np.int32))
ocv = O.Convex(10, -9, 9, -20, -2, -40, -1, -100000)
for i, rd in enumerate(reads):
    w = O.convex_align_pair(refs[0], rd, ocv)
    if w["score"] != int(br.score_scaled[i]) or O.cigar_str(w["cigar"]) != br.cigar_string(i):
        bad += 1
print("convex ok")
al.close()
print("mismatches", bad)
sys.exit(1 if bad else 0)
