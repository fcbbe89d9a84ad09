"""Randomised GPU-vs-oracle sweep over kernel families, modes and geometries (run on the GPU box for a fixed time budget):
   python tools/fuzz_gpu.py [seconds] [seed] [iterations to replay]      (CLQ_FUZZ_WIDE=1: more option and length draws)"""
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _oracle as O
from clique_b200 import AffineScoring, Aligner, Reference, ReferenceManager, RustBioScoring, TwoPieceScoring
from clique_b200.aligner import pack_reads

def reset():
    for k, v in (("force_cfg", -1), ("no_pack", 0), ("max_scratch_bytes", 40 << 30), ("no_group", 0), ("force_generic", 0), ("no_madd", 0), ("no_adapt", 0), ("no_long8", 0), ("no_overlap", 0)):
        al.set_option(k, v)


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
wide = os.environ.get("CLQ_FUZZ_WIDE", "0") == "1"   # also draw scratch sub-batching, no_group, force_generic and 2.6 / 5.2 kb lengths
# CLQ_FUZZ_MODES=fixed,exhaustive,quick restricts the drawn modes; CLQ_FUZZ_NO_PACK_P=0 keeps every batch on the s16x2 kernels when it fits
MODES = os.environ.get("CLQ_FUZZ_MODES", "fixed,fixed,exhaustive,quick,rustbio,convex,bandk").split(",")
NO_PACK_P = float(os.environ.get("CLQ_FUZZ_NO_PACK_P", "0.3"))
only = set(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else None   # replay only these iterations (same RNG stream)
rng = np.random.default_rng(seed)
SC = [(10.0, -9.0, 9.0, -20.0, -2.0, 1.0), (5.0, -4.0, 4.0, -10.0, -0.5, 0.5), (10.0, -5.0, 8.0, -15.0, -1.0, 0.25), (6.0, -6.0, 5.0, -10.0, -10.0, 1.0),
      (1.0, -1.0, 1.0, -5.0, -1.0, 1.0)]


def rs(n, a=b"ACGT"):
    return bytes(rng.choice(list(a), size=n).astype(np.uint8)) if n else b""


def mut(s, p):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < p / 3:
            continue
        if r < 2 * p / 3:
            out.append(int(rng.choice(list(b"ACGT")))); continue
        if r < p:
            out += rs(int(rng.integers(1, 5)))
        out.append(c)
    return bytes(out)


al = Aligner(device=0, max_reads=4096, max_read_bytes=1 << 24, max_read_len=1 << 13, cigar_ops_per_read=512, n_slots=2)
t_end, it, bad, n_checked = time.time() + budget, 0, 0, 0
while time.time() < t_end:
    it += 1
    nref = int(rng.choice([1, 1, 2, 5]))
    alpha = [b"ACGT", b"ACGTN", b"ACGTN012", b"ACGTacgtN"][int(rng.integers(0, 4))]
    lmax = int(rng.choice([40, 150, 330, 700, 1300] + ([2600, 5200] if wide else [])))
    uniform = rng.random() < 0.4
    L0 = int(rng.integers(1, lmax))
    refs = [rs(L0 if uniform else int(rng.integers(1, lmax)), alpha) for _ in range(nref)]
    n = int(rng.integers(1, 260 if lmax <= 1300 else 24))
    reads, fixed = [], []
    for _ in range(n):
        k = int(rng.integers(0, nref))
        base = refs[k].replace(b"0", b"A").replace(b"1", b"C").replace(b"2", b"G")
        rd = mut(base, float(rng.choice([0.0, 0.03, 0.15, 0.4]))) if rng.random() < 0.85 else rs(int(rng.integers(0, lmax + 50)), b"ACGTN")
        if uniform:
            rd = (rd + rs(L0 + 8))[:L0 + 3]
        reads.append(rd); fixed.append(k)
    fixed = np.array(fixed, np.int32)
    al.set_references(ReferenceManager([Reference(r, b"r%d" % i) for i, r in enumerate(refs)]))
    qb, qo = pack_reads(reads)
    rb, ro = O.pack_seqs(refs)
    mode = str(rng.choice(MODES))
    sc = SC[int(rng.integers(0, len(SC)))]
    cfg = int(rng.choice([-1, -1, 0, 1, 2, 3, 4, 5]))
    al.set_option("force_cfg", cfg if mode != "convex" else min(cfg, 4))
    wopts = {"no_pack": int(rng.random() < NO_PACK_P)}
    if wide:  # sub-batches of the direction-bit scratch, int32 multi-reference traceback, generic kernels
        wopts.update({"max_scratch_bytes": int(rng.choice([40 << 30, 1 << 20, 16 << 20])), "no_group": int(rng.random() < 0.3),
                      "force_generic": int(rng.random() < 0.15), "no_madd": int(rng.random() < 0.3), "no_adapt": int(rng.random() < 0.2),
                      "no_long8": int(rng.random() < 0.3), "no_overlap": int(rng.random() < 0.3)})
    for k_, v_ in wopts.items():
        al.set_option(k_, v_)
    tags = bool(rng.random() < 0.5) and mode not in ("convex",)
    ctx = (it, mode, sc, cfg, nref, n, lmax, uniform)
    if only is not None and it not in only:
        if mode not in ("rustbio", "convex"):   # consume the same draws as the real branch
            rng.choice(["readlen", "maxlen"])
            if mode == "bandk":
                rng.choice([1, 3, 10, 50, 400])
        reset()
        if it > max(only):
            break
        continue
    first = True
    p2 = ((it * 2654435761) >> 7) % 3 == 0   # every third batch or so ships 2-bit packed (clq_submit_packed2); no draw from rng
    if os.environ.get("CLQ_FUZZ_VERBOSE"):   # the context of every batch before it runs (to locate a crash)
        print("batch", ctx, "reflens", [len(r) for r in refs], "readlens", [len(r) for r in reads][:30], "opts", wopts, flush=True)
    try:
        if mode == "rustbio":
            br = al.align_batch(qb, qo, RustBioScoring(), "fixed", "maxlen", fixed_ref=fixed, extract_tags=tags, packed2=p2)
            for i, rd in enumerate(reads):
                w = O.rustbio_global(refs[fixed[i]], rd)
                st = int(br.status[i])
                if st == 2:
                    continue  # a read byte without a class column: refused, not mis-scored
                if st != 0 or int(br.score_scaled[i]) != w["score"] or O.cigar_str(br.cigar(i)) != O.cigar_str(w["cigar"]):
                    bad += 1
                    if first:
                        first = False
                        print("MISMATCH", ctx, "read", i, st, int(br.score_scaled[i]), w["score"], flush=True)
                n_checked += 1
        elif mode == "convex":
            if any(b in alpha for b in b"012acgt"):
                continue
            cv = TwoPieceScoring(10, -9, 9, -20, -2, -40, -1)
            ocv = O.Convex(10, -9, 9, -20, -2, -40, -1, -100000)
            br = al.align_batch(qb, qo, cv, "fixed", "readlen", fixed_ref=fixed, packed2=p2)
            for i, rd in enumerate(reads):
                w = O.convex_align_pair(refs[fixed[i]], rd, ocv)
                if int(br.status[i]) != 0 or int(br.score_scaled[i]) != w["score"] or br.cigar_string(i) != O.cigar_str(w["cigar"]):
                    bad += 1
                    if first:
                        first = False
                        print("MISMATCH", ctx, "read", i, int(br.status[i]), int(br.score_scaled[i]), w["score"], flush=True)
                n_checked += 1
        else:
            band = str(rng.choice(["readlen", "maxlen"]))
            search = mode if mode in ("exhaustive", "quick") else "fixed"
            if mode == "bandk":
                band = int(rng.choice([1, 3, 10, 50, 400]))
            br = al.align_batch(qb, qo, AffineScoring(*sc), search, band, fixed_ref=fixed if search == "fixed" else None, extract_tags=tags, packed2=p2)
            want = O.align_batch(rb, ro, qb, qo, sc, search=search, fixed_ref=fixed if search == "fixed" else None,
                                 band_mode="k" if mode == "bandk" else band, band_k=band if mode == "bandk" else 0, threads=16,
                                 traceback_all=False)
            for i, rd in enumerate(reads):
                ws = int(want["status"][i])
                ok = int(br.status[i]) == ws
                if ws != 5:
                    ok = ok and int(br.ref_index[i]) == int(want["ref_index"][i]) and int(br.score_scaled[i]) == want["score"][i] * br.scale
                if ws == 0:
                    o, l = int(want["cigar_off"][i]), int(want["cigar_len"][i])
                    ok = ok and O.cigar_str(br.cigar(i)) == O.cigar_str(want["cigar_pool"][o:o + l])
                    ok = ok and (int(br.matches[i]), int(br.mismatches[i])) == (int(want["matches"][i]), int(want["mismatches"][i]))
                    # extract_tagged_sequences numbers its upper-case regions in a u8 ('A' + region): past ~190 regions the counter
                    # wraps into '0'..'9' and the oracle's digit keys are no longer the digit columns -- not a case the GPU tags model.
                    if ok and tags and len(re.findall(rb"[A-Z]+", refs[int(want["ref_index"][i])])) < 180:
                        ri = int(want["ref_index"][i])
                        ra, qa = O.apply_cigar(refs[ri], rd, want["cigar_pool"][o:o + l])
                        exp = {k: v for k, v in O.extract_tagged_sequences(qa, ra).items() if 48 <= k <= 57}
                        ok = br.tag_strings(i, refs[ri]) == exp
                if not ok:
                    bad += 1
                    if first:
                        first = False
                        o, l = int(want["cigar_off"][i]), int(want["cigar_len"][i])
                        if ws == 0:
                            ri = int(want["ref_index"][i])
                            ra, qa = O.apply_cigar(refs[ri], rd, want["cigar_pool"][o:o + l])
                            exp = {k: v for k, v in O.extract_tagged_sequences(qa, ra).items() if 48 <= k <= 57}
                            got = br.tag_strings(i, refs[ri]) if tags else None
                            print("  rm gpu", int(br.matches[i]), int(br.mismatches[i]), "oracle", int(want["matches"][i]), int(want["mismatches"][i]),
                                  "tags equal", got == exp, "alpha", alpha, "tag_stride", None if br.tags is None else br.tags.shape,
                                  "exp", {k: v[:12] for k, v in exp.items()}, "got", None if got is None else {k: v[:12] for k, v in got.items()}, flush=True)
                        print("MISMATCH", ctx, "band", band, "tags", tags, "read", i, "len", len(rd), "reflens", [len(r) for r in refs],
                              "gpu", int(br.status[i]), int(br.ref_index[i]), int(br.score_scaled[i]) / br.scale, O.cigar_str(br.cigar(i))[:60],
                              "oracle", ws, int(want["ref_index"][i]), want["score"][i], O.cigar_str(want["cigar_pool"][o:o + l])[:60], flush=True)
                n_checked += 1
    finally:
        reset()
al.close()
print("iterations", it, "checked pairs", n_checked, "mismatches", bad)
# debug builds (-DCLQ_PACK_CANARY=1, tools/canary_gpu.sh): s16x2 tasks whose stored values were tracked / tasks that left [64, 32767]
# without being covered by the retry pass
import ctypes
from clique_b200 import load_library
_lib = load_library()
if hasattr(_lib, "clq_debug_canary"):
    chk, vio = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    if _lib.clq_debug_canary(ctypes.byref(chk), ctypes.byref(vio)) == 0:
        print("canary: s16x2 tasks checked", chk.value, "window violations", vio.value)
        bad += vio.value
sys.exit(1 if bad else 0)
