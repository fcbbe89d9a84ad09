"""Deterministic synthetic workloads for the bench configs of BASELINE.json / SURVEY.md section 8d.

Everything is generated with numpy's PCG64 from a fixed seed (0xC11C0000 + config id), so the GPU run, the oracle and
the CPU baseline see byte-identical inputs.  Amplicons and the 64-reference panel come from the reference's own test
data (tests/golden/reference_goldens.json, lifted by tests/golden/make_golden.py).
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_goldens.json")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
CLI_SCORING = (10.0, -9.0, 9.0, -20.0, -2.0, 1.0)  # alignment_functions.rs:104-111

_gold = None


def goldens():
    global _gold
    if _gold is None:
        with open(GOLDEN) as f:
            _gold = json.load(f)
    return _gold


def rand_bases(rng, n):
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def substitute(rng, seq, mask):
    """replace seq[mask] by one of the three other bases (non-ACGT bytes become a random base)"""
    idx = np.zeros(256, np.uint8)
    idx[ACGT] = np.arange(4, dtype=np.uint8)
    cur = idx[seq[mask]]
    seq[mask] = ACGT[(cur + rng.integers(1, 4, size=cur.size, dtype=np.uint8)) & 3]


def noisy_copies(rng, template, n, p_sub, p_ins, p_del, chunk_cells=1 << 24):
    """n noisy copies of `template` (uint8 array): per base, in order: delete (p_del), substitute (p_sub), and insert a
    geometric(0.5) run of random bases before it (p_ins).  Returns (flat bytes, uint64 offsets)."""
    L = len(template)
    outs, lens = [], []
    m_chunk = max(1, chunk_cells // max(L, 1))
    for lo in range(0, n, m_chunk):
        m = min(m_chunk, n - lo)
        dele = rng.random((m, L)) < p_del
        sub = (rng.random((m, L)) < p_sub) & ~dele
        ins = rng.random((m, L)) < p_ins
        ins_len = np.where(ins, rng.geometric(0.5, size=(m, L)), 0).astype(np.int64)
        cnt = ins_len + (~dele)
        rl = cnt.sum(axis=1)
        flat_cnt = cnt.ravel()
        ends = np.cumsum(flat_cnt)
        out = rand_bases(rng, int(ends[-1]) if ends.size else 0)
        base = np.tile(template, m).reshape(m, L).copy()
        substitute(rng, base, sub)
        keep = ~dele.ravel()
        pos = (ends - 1)[keep]  # the kept base sits after its inserted run
        out[pos] = base.ravel()[keep]
        outs.append(out)
        lens.append(rl)
    lens = np.concatenate(lens) if lens else np.zeros(0, np.int64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens, dtype=np.uint64)
    data = np.concatenate(outs) if outs else np.zeros(0, np.uint8)
    return data, off


def fix_length(rng, data, off, target):
    """truncate / pad (random bases) every read to its target length; returns (flat, offsets)"""
    n = len(off) - 1
    target = np.broadcast_to(np.asarray(target, dtype=np.int64), (n,))
    noff = np.zeros(n + 1, np.uint64)
    noff[1:] = np.cumsum(target, dtype=np.uint64)
    out = rand_bases(rng, int(noff[-1]))
    cur = (off[1:] - off[:-1]).astype(np.int64)
    take = np.minimum(cur, target)
    # flat gather of the first take[i] bytes of read i
    tot = int(take.sum())
    rid = np.repeat(np.arange(n), take)
    within = np.arange(tot) - np.repeat(np.cumsum(take) - take, take)
    out[noff[:-1].astype(np.int64)[rid] + within] = data[off[:-1].astype(np.int64)[rid] + within]
    return out, noff


def config_c2(n_reads, seed=0xC11C0002):
    """C2 'mouse lineage': 300 bp Illumina-like reads vs the 215 bp lineage amplicon (digits = UMI/tag positions)."""
    rng = np.random.default_rng(seed)
    tmpl = np.frombuffer(goldens()["amplicon_c2"].encode(), dtype=np.uint8)
    L = len(tmpl)
    tag = (tmpl < 58) | (tmpl == ord("N"))  # tag / UMI positions carry random bases in a read
    base = np.tile(tmpl, n_reads).reshape(n_reads, L).copy()
    base[:, tag] = rand_bases(rng, n_reads * int(tag.sum())).reshape(n_reads, -1)
    # one lineage deletion in half of the reads
    has_del = rng.random(n_reads) < 0.5
    start = rng.integers(60, 181, size=n_reads)
    dlen = np.where(has_del, np.minimum(rng.geometric(0.15, size=n_reads), L - start), 0)
    j = np.arange(300)[None, :]
    src = j + np.where(j >= start[:, None], dlen[:, None], 0)
    reads = rand_bases(rng, n_reads * 300).reshape(n_reads, 300)
    ok = src < L
    reads[ok] = np.take_along_axis(base, np.minimum(src, L - 1), axis=1)[ok]
    sub = rng.random((n_reads, 300)) < 0.003
    substitute(rng, reads, sub)
    # rare single-base indel errors (p = 1e-4 per base)
    for kind in (0, 1):
        hit = np.nonzero(rng.random(n_reads) < 300 * 1e-4)[0]
        pos = rng.integers(0, 300, size=hit.size)
        for i, p in zip(hit, pos):
            row = reads[i]
            if kind == 0:
                row[p:-1] = row[p + 1:].copy(); row[-1] = ACGT[rng.integers(0, 4)]
            else:
                row[p + 1:] = row[p:-1].copy(); row[p] = ACGT[rng.integers(0, 4)]
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(300))
    return {"name": "C2", "refs": [tmpl.tobytes()], "ref_names": [b"lineage_amplicon"], "read_bytes": reads.ravel(),
            "read_off": off, "fixed_ref": np.zeros(n_reads, np.int32), "scoring": CLI_SCORING, "search": "fixed",
            "band": "readlen", "cells": int(L) * 300 * n_reads}


def config_c3(n_reads, seed=0xC11C0003, unique=None):
    """C3 'ONT': ~1 kb reads (N(1000,120) clipped to [500,1900]), 4% ins + 4% del + 2% sub, vs a 1 kb amplicon."""
    rng = np.random.default_rng(seed)
    tmpl = np.frombuffer(goldens()["amplicon_c3"].encode(), dtype=np.uint8)
    nu = min(n_reads, unique or n_reads)
    data, off = noisy_copies(rng, tmpl, nu, 0.02, 0.04, 0.04)
    target = np.clip(np.rint(rng.normal(1000, 120, size=nu)), 500, 1900).astype(np.int64)
    data, off = fix_length(rng, data, off, target)
    data, off = _tile(data, off, n_reads)
    lens = (off[1:] - off[:-1]).astype(np.int64)
    return {"name": "C3", "refs": [tmpl.tobytes()], "ref_names": [b"ont_amplicon"], "read_bytes": data, "read_off": off,
            "fixed_ref": np.zeros(n_reads, np.int32), "scoring": CLI_SCORING, "search": "fixed", "band": "readlen",
            "cells": int(len(tmpl) * lens.sum())}


def config_c4(n_reads, seed=0xC11C0004, search="exhaustive", unique=None):
    """C4 panel: reads drawn from the first 64 references of 18guide1_pcr_sequence.fasta (302 bp each), Illumina errors."""
    rng = np.random.default_rng(seed)
    recs = goldens()["fastas"]["18guide1_pcr_sequence.first64"]
    refs = [r["seq"].encode() for r in recs]
    nu = min(n_reads, unique or n_reads)
    which = rng.integers(0, len(refs), size=nu)
    mat = np.stack([np.frombuffer(r, dtype=np.uint8) for r in refs])
    reads = mat[which].copy()
    nmask = reads == ord("N")
    reads[nmask] = rand_bases(rng, int(nmask.sum()))
    substitute(rng, reads, rng.random(reads.shape) < 0.003)
    off = np.arange(nu + 1, dtype=np.uint64) * np.uint64(reads.shape[1])
    data, off = _tile(reads.ravel(), off, n_reads)
    truth = np.resize(which, n_reads).astype(np.int32)
    L2 = reads.shape[1]
    return {"name": "C4", "refs": refs, "ref_names": [r["name"].encode() for r in recs], "read_bytes": data, "read_off": off,
            "fixed_ref": None, "truth": truth, "scoring": CLI_SCORING, "search": search, "band": "readlen",
            "cells": int(sum(len(r) for r in refs)) * L2 * n_reads if search == "exhaustive" else None}


C5_LENGTHS = [300, 450, 700, 1000, 1500, 2200, 3300, 5000]


def config_c5(n_reads, seed=0xC11C0005, unique_per_amplicon=256):
    """C5 mixed 300 bp - 5 kb: eight random amplicons, every read a noisy full-length copy of one of them."""
    rng = np.random.default_rng(seed)
    refs = [rand_bases(rng, L) for L in C5_LENGTHS]
    pools = []
    for t in refs:
        if len(t) <= 700:
            pools.append(noisy_copies(rng, t, unique_per_amplicon, 0.003, 0.0001, 0.0001))
        else:
            pools.append(noisy_copies(rng, t, unique_per_amplicon, 0.02, 0.04, 0.04))
    which = rng.integers(0, len(refs), size=n_reads)
    pick = rng.integers(0, unique_per_amplicon, size=n_reads)
    lens = np.array([int(pools[w][1][p + 1] - pools[w][1][p]) for w, p in zip(which, pick)], dtype=np.int64)
    off = np.zeros(n_reads + 1, np.uint64)
    off[1:] = np.cumsum(lens, dtype=np.uint64)
    data = np.empty(int(off[-1]), np.uint8)
    for i, (w, p) in enumerate(zip(which, pick)):
        d, o = pools[w]
        data[int(off[i]):int(off[i + 1])] = d[int(o[p]):int(o[p + 1])]
    cells = int(sum(len(refs[w]) * l for w, l in zip(which, lens)))
    return {"name": "C5", "refs": [r.tobytes() for r in refs], "ref_names": [b"amp%d" % L for L in C5_LENGTHS], "read_bytes": data,
            "read_off": off, "fixed_ref": which.astype(np.int32), "scoring": CLI_SCORING, "search": "fixed", "band": "readlen",
            "cells": cells}


def _tile(data, off, n):
    """repeat a pool of unique reads cyclically up to n reads"""
    nu = len(off) - 1
    if n <= nu:
        return data[:int(off[n])], off[:n + 1].copy()
    reps = -(-n // nu)
    lens = np.tile((off[1:] - off[:-1]).astype(np.int64), reps)[:n]
    noff = np.zeros(n + 1, np.uint64)
    noff[1:] = np.cumsum(lens, dtype=np.uint64)
    full = np.tile(data[:int(off[-1])], reps)[:int(noff[-1])]
    return full, noff


CONFIGS = {"C2": config_c2, "C3": config_c3, "C4": config_c4, "C5": config_c5}
