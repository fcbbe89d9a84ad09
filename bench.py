#!/usr/bin/env python3
"""bench.py -- reads/s and GCUPS of the batched amplicon-alignment hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4|C5] [--reads R]

A step is one pass of the hot path (Gotoh fill + traceback + CIGAR, bit-exact vs the reference) over one batch of
synthetic reads per GPU.  Default workload = BASELINE.json configs[1] ("mouse_lineage_test: 1M synthetic ~300 bp
Illumina-like reads vs single lineage amplicon, affine gap"), 1M reads per GPU per step (weak scaling, reads are
independent: read-sharded, no data-path collective).

  value  : whole-job reads/s with the batch already resident in HBM (kernel launches only, CUDA-event timed)
  e2e    : the same through clq_submit/clq_wait with pinned HOST buffers, H2D/D2H inside the timed region
  roofline / cpu_baseline : see DESIGN.md section "Measurement"
`--impl reference` times the CPU restatement of the reference's aligner (oracle/, kind "port": the Rust crate cannot be
built in this image) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CELL_TB = 18   # algorithmic INT32 ops per cell with traceback bits
OPS_PER_CELL_SCORE = 12
WORKLOAD_DESC = {
    "C2": "C2 mouse_lineage: 300 bp Illumina-like reads vs the 215 bp lineage amplicon, affine gap 10/-9/9/-20/-2, CIGAR out",
    "C3": "C3 ONT: ~1 kb reads (8% indel) vs a 1 kb amplicon, affine gap, CIGAR out",
    "C4": "C4 panel: 302 bp reads, best-candidate selection across 64 reference amplicons (exhaustive), CIGAR out",
    "C5": "C5 mixed 300 bp-5 kb reads vs 8 amplicons, affine gap, CIGAR out",
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2]); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def int32_peak(live=True):
    """Measured INT32 ALU-pipe issue rate (T lane-ops/s): run tools/int_peak here, else the committed round-1 measurement."""
    exe = os.path.join(ROOT, "tools", "int_peak")
    rows, how = [], None
    if live and os.path.exists(exe):
        try:
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            rows = [json.loads(l) for l in out.splitlines() if l.startswith("{") and '"op"' in l]
            how = "measured live (tools/int_peak: VIADDMNMX/VIMNMX3/PRMT issue rate on this GPU)"
        except Exception:
            rows = []
    if not rows:
        p = os.path.join(ROOT, "profiles", "int_peak_r01.jsonl")
        rows = [json.loads(l) for l in open(p) if '"op"' in l]
        how = "fallback: profiles/int_peak_r01.jsonl (earlier measurement on this pool)"
    alu = [r["tops"] for r in rows if r["op"] in ("viaddmax_s32", "vimax3_s32", "prmt", "lop3")]
    return max(alu), how


def make_workload(name, n_reads, search=""):
    from clique_b200 import synth
    if name == "C2":
        return synth.config_c2(n_reads)
    if name == "C3":
        return synth.config_c3(n_reads, unique=min(n_reads, 65536))
    if name == "C4":
        return synth.config_c4(n_reads, search=search or "exhaustive", unique=min(n_reads, 262144))
    return synth.config_c5(n_reads)


def default_reads(name):
    return {"C2": 1_000_000, "C3": 100_000, "C4": 100_000, "C5": 60_000}[name]


def cpu_reference_run(c, n_sample, threads):
    """time the oracle port (faithful f64 restatement, fill + traceback per candidate) on a prefix of the workload"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    rb, ro = O.pack_seqs(c["refs"])
    off = c["read_off"][:n_sample + 1]
    qb = c["read_bytes"][:int(off[-1])]
    fr = None if c["fixed_ref"] is None else c["fixed_ref"][:n_sample]
    t0 = time.perf_counter()
    out = O.align_batch(rb, ro, qb, off, c["scoring"], search=c["search"], fixed_ref=fr, band_mode=c["band"], threads=threads,
                        traceback_all=True)
    dt = time.perf_counter() - t0
    return dt, out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    c = make_workload(args.workload, 20000 if args.workload == "C2" else 2000, args.search)
    n_have = len(c["read_off"]) - 1
    probe = min(n_have, 256 if args.workload == "C2" else 16)
    dt, out = cpu_reference_run(c, probe, threads)
    per_read = dt / probe
    n_sample = int(max(probe, min(n_have, args.ref_step_seconds / per_read)))
    times, cells = [], 0
    for it in range(args.warmup + args.steps):
        dt, out = cpu_reference_run(c, n_sample, threads)
        if it >= args.warmup:
            times.append(dt); cells = int(out["cells"])
    ms = 1e3 * float(np.mean(times))
    value = n_sample / (ms / 1e3)
    sample = "first %d reads of the %s workload per step (seed-identical inputs), %d threads, f64 port of the reference aligner" % (n_sample, args.workload, threads)
    line = {"impl": "reference", "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gcups": cells / (ms / 1e3) / 1e9,
            "config": {"workload": WORKLOAD_DESC[args.workload], "reads_per_step": n_sample},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


KERNEL_NAMES = {0: "generic int32", 1: "FAST int32 (PRMT profile + DPX)", 3: "PACK s16x2 (two reads per lane group, DPX)",
                35: "PACK s16x2 with static row slope (two reads per lane group, DPX; M step off the ALU pipe)",
                67: "PACK s16x2 with adaptive row bias + int32 retry pass (two reads per lane group, DPX)",
                5: "CONVEX int32 (two-piece affine, DPX)", 7: "CONVEX PACK s16x2 (two-piece affine, two reads per lane group, DPX)"}


def kernel_name(variant):
    k = KERNEL_NAMES.get(variant & 103, KERNEL_NAMES.get(variant & 7, "variant %d" % (variant & 127)))
    if variant & 16:
        k += " [rust-bio global semantics]"
    return k


def rotate_for_rank(c, n, rank):
    """weak scaling: every rank aligns its own n reads per step (same generator, rank-specific rotation of the batch)"""
    if not rank:
        return
    k = (rank * 7919) % n
    lens = (c["read_off"][1:] - c["read_off"][:-1]).astype(np.int64)
    if (lens == lens[0]).all():
        L = int(lens[0])
        c["read_bytes"] = np.roll(c["read_bytes"].reshape(n, L), k, axis=0).ravel()
        if c["fixed_ref"] is not None:
            c["fixed_ref"] = np.roll(c["fixed_ref"], k)


def algorithmic_ops(workload, convex, cells, c, total_bytes):
    """12 INT32 ops per score-only cell, 18 with direction bits (two-piece affine: 20 / 30)"""
    if workload == "C4":   # n score-only fills + 1 traceback fill per read (cells counts both)
        tb_cells = float(sum(len(r) for r in c["refs"])) / len(c["refs"]) * float(total_bytes)
        return cells * OPS_PER_CELL_SCORE + min(tb_cells, cells) * (OPS_PER_CELL_TB - OPS_PER_CELL_SCORE), "12 score-only + 18 traceback"
    ops = 30 if convex else OPS_PER_CELL_TB
    return cells * ops, ops


def oracle_parity(c, res, ns, convex=False, rustbio=False, threads=None):
    """spot check of the timed batch against the CPU oracle (the checker, never the thing measured)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    bad = 0
    if convex:
        ocv = O.Convex(10, -9, 9, -20, -2, -40, -1, -100000)
        for i in range(ns):
            rd = bytes(c["read_bytes"][int(c["read_off"][i]):int(c["read_off"][i + 1])])
            w = O.convex_align_pair(c["refs"][int(c["fixed_ref"][i]) if c["fixed_ref"] is not None else int(res.ref_index[i])], rd, ocv)
            if w["score"] != int(res.score_scaled[i]) or O.cigar_str(w["cigar"]) != res.cigar_string(i):
                bad += 1
        return {"checked_reads": ns, "mismatches": bad, "oracle": "orc_convex_align_pair (self-pinned)"}
    if rustbio:
        for i in range(ns):
            rd = bytes(c["read_bytes"][int(c["read_off"][i]):int(c["read_off"][i + 1])])
            w = O.rustbio_global(c["refs"][int(c["fixed_ref"][i])], rd)
            if w["score"] != int(res.score_scaled[i]) or O.cigar_str(w["cigar"]) != res.cigar_string(i):
                bad += 1
        return {"checked_reads": ns, "mismatches": bad, "oracle": "orc_rustbio_global (parity unpinned)"}
    dt, out = cpu_reference_run(c, ns, threads or (os.cpu_count() or 1))
    for i in range(ns):
        o, l = int(out["cigar_off"][i]), int(out["cigar_len"][i])
        if int(res.score_scaled[i]) != out["score"][i] * res.scale or not np.array_equal(res.cigar(i), out["cigar_pool"][o:o + l]) \
                or int(res.ref_index[i]) != int(out["ref_index"][i]):
            bad += 1
    return {"checked_reads": ns, "mismatches": bad, "oracle_seconds": dt, "oracle_cells": int(out["cells"])}


def run_config(args, workload, n, search, convex, rustbio, steps, warmup, e2e_chunks, rank, local_rank, world, barrier, clock_device=None,
               packed2_e2e=False):
    """Device-resident and end-to-end passes of one workload on this rank's GPU; max over ranks of the timings."""
    import torch
    import torch.distributed as dist
    from clique_b200 import AffineScoring, Aligner, Reference, ReferenceManager, RustBioScoring, TwoPieceScoring
    dev = local_rank if world > 1 else 0
    c = make_workload(workload, n, search)
    rotate_for_rank(c, n, rank)
    total_bytes = int(c["read_off"][-1])
    ops_per_read = {"C2": 12, "C3": 400, "C4": 12, "C5": 700}[workload]
    al = Aligner(device=dev, max_reads=n, max_read_bytes=total_bytes + 64, max_read_len=1 << 15, max_refs=max(64, len(c["refs"])),
                 cigar_ops_per_read=ops_per_read, n_slots=2)
    al.set_references(ReferenceManager([Reference(r, nm) for r, nm in zip(c["refs"], c["ref_names"])]))
    sc = TwoPieceScoring(10, -9, 9, -20, -2, -40, -1) if convex else AffineScoring(*c["scoring"])
    sci = sc.to_int()
    if rustbio:
        assert c["search"] == "fixed" and not convex, "--rustbio is the single-reference branch"
        sci = RustBioScoring()   # Aligner.launch adds CLQ_RUSTBIO for this scoring type
        c["band"] = "maxlen"
    for opt in ("force_cfg", "force_generic", "debug_flags", "no_pack", "no_madd", "no_adapt", "no_long8", "no_overlap", "max_scratch_bytes"):     # experiment knobs, e.g. CLQ_FORCE_CFG=3
        if os.environ.get("CLQ_" + opt.upper()):
            al.set_option(opt, int(os.environ["CLQ_" + opt.upper()]))
    score_only = bool(int(os.environ.get("CLQ_SCORE_ONLY", "0")))

    # pinned host buffers for the e2e path, split into chunks that alternate over the two stream slots
    h_bytes = al.alloc_pinned(total_bytes, np.uint8)
    h_bytes[:] = c["read_bytes"][:total_bytes]
    nch = max(1, e2e_chunks)
    bounds = [n * i // nch for i in range(nch + 1)]
    chunks = []
    for i in range(nch):
        lo, hi = bounds[i], bounds[i + 1]
        off = al.alloc_pinned(hi - lo + 1, np.uint64)
        off[:] = c["read_off"][lo:hi + 1] - c["read_off"][lo]
        fr = None
        if c["fixed_ref"] is not None:
            fr = al.alloc_pinned(hi - lo, np.int32)
            fr[:] = c["fixed_ref"][lo:hi]
        chunks.append((h_bytes[int(c["read_off"][lo]):int(c["read_off"][hi])], off, fr))

    # ---------------- device-resident throughput: inputs in HBM, launches only ----------------
    al.upload(0, h_bytes, c["read_off"], c["fixed_ref"])
    al.sync(0)
    for _ in range(warmup):
        al.launch(0, sci, c["search"], c["band"], score_only)
        al.sync(0)
    clocks = ClockSampler(dev) if clock_device is not None else None
    if clocks:
        clocks.start()
    barrier()
    step_ms, dp_ms, launches, cells, variant, sub_batches = [], [], 0, 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        al.launch(0, sci, c["search"], c["band"], score_only)
        st = al.stats(0)       # synchronises the slot's stream; times come from CUDA events on that stream
        step_ms.append(st["kernel_ms"]); dp_ms.append(st["dp_ms"]); launches += st["launches"]; cells = st["cells"]; variant = st["variant"]
        sub_batches = st["sub_batches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk = clocks.stop() if clocks else None
    res = al.wait(0, copy=True)
    pack_retries = al.stats(0).get("pack_retries", 0)   # set by clq_wait: reads the adaptive s16x2 kernel handed to the int32 retry pass
    n_ok = int((res.status == 0).sum())
    dev_ms_total = float(np.sum(step_ms))
    t = torch.tensor([dev_ms_total, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_total, wall_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms_total / steps

    # ---------------- end to end: pinned host buffers -> clq_submit -> clq_wait, copies inside the timed region ----------------
    def e2e_step(collect=False):
        h2d = d2h = 0
        pending = []
        last = None
        for i, (rb, off, fr) in enumerate(chunks):
            slot = i % 2
            if len(pending) == 2:
                s0 = pending.pop(0)
                last = al.wait(s0, copy=False)
                if collect:
                    stt = al.stats(s0); h2d += stt["h2d_bytes"]; d2h += stt["d2h_bytes"]
            al.submit(slot, rb, off, sci, c["search"], c["band"], fixed_ref=fr)
            pending.append(slot)
        for s0 in pending:
            last = al.wait(s0, copy=False)
            if collect:
                stt = al.stats(s0); h2d += stt["h2d_bytes"]; d2h += stt["d2h_bytes"]
        return h2d, d2h, int(last.score_scaled[0])

    h2d, d2h, _chk = e2e_step(collect=True)
    for _ in range(warmup):
        e2e_step()

    def e2e_run(k_steps):
        """K steps as the product runs them: one double-buffered stream of chunks (chunk k+1 uploads while chunk k computes,
        also across step boundaries); every chunk's inputs are copied from pinned host memory and its results read back."""
        busy, k, last = [False, False], 0, None
        for _step in range(k_steps):
            for rb, off, fr in chunks:
                s0 = k % 2
                if busy[s0]:
                    last = al.wait(s0, copy=False)
                al.submit(s0, rb, off, sci, c["search"], c["band"], fixed_ref=fr)
                busy[s0] = True
                k += 1
        for s0 in (k % 2, (k + 1) % 2):
            if busy[s0]:
                last = al.wait(s0, copy=False)
        return int(last.score_scaled[0])

    barrier()
    t0 = time.perf_counter()
    e2e_run(steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0]) / steps

    # ---------------- the same stream with the chunks 2-bit packed (clq_submit_packed2): a quarter of the read bytes on the wire ----------------
    packed = None
    if packed2_e2e:
        plain_chunks = chunks
        t0 = time.perf_counter()
        pchunks, n_exc = [], 0
        for rb, off, fr in plain_chunks:
            words = al.alloc_pinned((len(rb) + 15) // 16 + 4, np.uint32)
            pk = al.pack_reads(rb, len(rb), out_words=words)
            n_exc += len(pk.exc_pos)
            pchunks.append((pk, off, fr))
        pack_s = time.perf_counter() - t0          # host packing pass, one thread (reported, not inside the packed timer)
        chunks = pchunks
        h2d_p, d2h_p, _ = e2e_step(collect=True)
        for _ in range(warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_run(steps)
        barrier()
        p_ms = 1e3 * (time.perf_counter() - t0)
        t = torch.tensor([p_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # results of the packed route against the device-resident ASCII launch above, record for record
        al.submit(0, pchunks[0][0], pchunks[0][1], sci, c["search"], c["band"], fixed_ref=pchunks[0][2])
        rp = al.wait(0, copy=True)
        k0 = len(pchunks[0][1]) - 1
        same = bool((rp.score_scaled == res.score_scaled[:k0]).all() and (rp.cigar_len == res.cigar_len[:k0]).all()
                    and (rp.status == res.status[:k0]).all() and (rp.ref_index == res.ref_index[:k0]).all()
                    and all(np.array_equal(rp.cigar(i), res.cigar(i)) for i in range(0, k0, max(1, k0 // 2000))))
        packed = {"ms": float(t[0]) / steps, "h2d": int(h2d_p), "d2h": int(d2h_p), "exceptions": int(n_exc), "same": same,
                  "host_pack_gb_s_one_thread": total_bytes / pack_s / 1e9}
        chunks = plain_chunks
    al.close()
    return {"packed2": packed, "c": c, "n": n, "total_bytes": total_bytes, "res": res, "n_ok": n_ok, "ms_per_step": ms_per_step, "wall_ms": wall_ms,
            "dp_ms": float(np.mean(dp_ms)), "launches": int(launches), "cells": int(cells), "variant": int(variant), "clk": clk,
            "e2e_ms": e2e_ms, "h2d": int(h2d), "d2h": int(d2h), "nch": nch, "pack_retries": int(pack_retries), "sub_batches": int(sub_batches)}


def api_pass(c, devices, n_reads_total, passes=3, batch_reads=1 << 18, fillers=2, repeat=1):
    """The product's batch loop (C++ clique::ShardedAligner::align_reads_span, the analogue of align_reads' par_bridge loop,
    alignment_functions.rs:135) over `devices`, fed from plain UNPINNED host memory: staging copy into page-locked batches, H2D,
    kernels, D2H and the copy of the records into the caller's arrays are all inside the timed region (C++ steady_clock around
    the loop; the first pass page-locks the staging buffers and is not the one reported)."""
    from clique_b200 import AffineScoring
    from clique_b200.host import align_reads_span
    lens = (c["read_off"][1:] - c["read_off"][:-1]).astype(np.int64)
    mean_len = float(lens.mean()) if len(lens) else 1.0
    ops_per_read = int(max(16, 12 if mean_len < 400 else mean_len * 0.7))
    batch_bytes = int(min(1 << 30, max(1 << 22, batch_reads * mean_len * 1.25)))
    rbytes, roff, fixed = c["read_bytes"][:int(c["read_off"][-1])], c["read_off"], c["fixed_ref"]
    if repeat > 1:   # one longer stream: the same reads `repeat` times over (the pipeline's fill / drain latency is paid once per stream)
        n0, tot = len(roff) - 1, int(roff[-1])
        rbytes = np.tile(rbytes, repeat)
        roff = np.concatenate([roff[:-1].astype(np.uint64) + np.uint64(tot * k) for k in range(repeat)] + [np.array([tot * repeat], np.uint64)])
        fixed = None if fixed is None else np.tile(fixed, repeat)
    br, st = align_reads_span(devices, c["refs"], rbytes, roff, AffineScoring(*c["scoring"]),
                              fixed_ref=fixed if len(c["refs"]) > 1 else None, batch_reads=batch_reads, batch_bytes=batch_bytes,
                              max_read_len=1 << 15, cigar_ops_per_read=ops_per_read, n_slots=2, fillers_per_device=fillers,
                              fast_lookup=(c["search"] == "quick"), passes=passes)
    return br, st


def sorted_c5_stream(n):
    """one C5 stream, shortest reads first: the worst case for a contiguous byte split across GPUs"""
    from clique_b200 import synth
    c = synth.config_c5(n)
    off = c["read_off"]
    lens = (off[1:] - off[:-1]).astype(np.int64)
    order = np.argsort(lens, kind="stable")
    noff = np.zeros(n + 1, np.uint64)
    noff[1:] = np.cumsum(lens[order], dtype=np.uint64)
    data = np.empty(int(noff[-1]), np.uint8)
    src0 = off[:-1].astype(np.int64)[order]
    rid = np.repeat(np.arange(n), lens[order])
    within = np.arange(int(noff[-1])) - np.repeat(noff[:-1].astype(np.int64), lens[order])
    data[:] = c["read_bytes"][src0[rid] + within]
    c = dict(c)
    c["read_bytes"], c["read_off"], c["fixed_ref"] = data, noff, c["fixed_ref"][order].astype(np.int32)
    return c


# the other north_star configs (BASELINE.json configs[2..4]) as short passes inside the same bench line: (key, workload, search,
# convex, reads per GPU per step, oracle-checked reads)
EXTRA_CONFIGS = [
    ("C3", "C3", "", False, 100_000, 96),
    ("C3_convex", "C3", "", True, 40_000, 48),
    ("C4_exhaustive", "C4", "exhaustive", False, 50_000, 48),
    ("C4_quick", "C4", "quick", False, 400_000, 2048),
    ("C5", "C5", "", False, 60_000, 24),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOAD_DESC))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: the config's size, 1M for C2)")
    ap.add_argument("--e2e-chunks", type=int, default=2)
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    ap.add_argument("--search", default="", choices=["", "exhaustive", "quick"], help="C4 only: candidate search (default exhaustive)")
    ap.add_argument("--convex", action="store_true", help="two-piece affine gaps o1=-20,e1=-2,o2=-40,e2=-1 (self-pinned semantics)")
    ap.add_argument("--rustbio", action="store_true", help="C2 only: the reference's current single-reference branch (rust-bio global 1/-1/-5/-1; parity unpinned)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-live-peak", action="store_true", help="use the committed INT32 peak instead of running tools/int_peak")
    ap.add_argument("--no-extra", action="store_true", help="skip the short passes of the other north_star configs (C3, C3 convex, C4 x2, C5)")
    ap.add_argument("--extra-steps", type=int, default=3)
    ap.add_argument("--no-api", action="store_true", help="skip the C++ align_reads loop measurements (e2e_api, sharded)")
    ap.add_argument("--sharded-reads-per-gpu", type=int, default=30_000)
    ap.add_argument("--no-packed2", action="store_true", help="skip the 2-bit packed upload measurements (e2e_packed2, e2e_api_packed2)")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits: no NCCL kernel may spin on a GPU another process is measuring
    n_gpus = world
    dev = local_rank if world > 1 else 0
    n = args.reads or default_reads(args.workload)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r = run_config(args, args.workload, n, args.search, args.convex, args.rustbio, args.steps, args.warmup, args.e2e_chunks,
                   rank, local_rank, world, barrier, clock_device=dev, packed2_e2e=(not args.no_packed2 and world == 1))   # side measurement: N = 1 only
    c, res, cells, variant, total_bytes = r["c"], r["res"], r["cells"], r["variant"], r["total_bytes"]
    ms_per_step, e2e_ms = r["ms_per_step"], r["e2e_ms"]
    value = n * n_gpus / (ms_per_step / 1e3)
    gcups = cells * n_gpus / (ms_per_step / 1e3) / 1e9
    e2e_value = n * n_gpus / (e2e_ms / 1e3)
    clk = r["clk"]

    peak, peak_how = (None, None)
    if rank == 0:
        peak, peak_how = int32_peak(not args.no_live_peak)

    # ---------------- the other north_star configs, short passes (every rank runs them; rank 0 checks parity at N=1) ----------------
    extras = {}
    if not args.no_extra and args.workload == "C2" and not args.convex and not args.rustbio:
        for key, wl, search, convex, n_x, n_chk in EXTRA_CONFIGS:
            try:
                x = run_config(args, wl, n_x, search, convex, False, args.extra_steps, 2, 2, rank, local_rank, world, barrier)
            except Exception as e:  # noqa: BLE001 -- a failing side config must not take the headline down; it is reported instead
                extras[key] = {"error": "%s: %s" % (type(e).__name__, e)}
                continue
            if rank == 0:
                alg_ops, ops = algorithmic_ops(wl, convex, x["cells"], x["c"], x["total_bytes"])
                achieved = alg_ops / (x["dp_ms"] / 1e3) / 1e12
                pack = 2 if (x["variant"] & 2) else 1
                ent = {"workload": WORKLOAD_DESC[wl].replace("(exhaustive)", "(%s)" % x["c"]["search"]) + (" [two-piece affine (convex) gaps, self-pinned]" if convex else ""),
                       "reads_per_gpu_per_step": n_x, "steps": args.extra_steps, "reads_s": n_x * n_gpus / (x["ms_per_step"] / 1e3),
                       "gcups": x["cells"] * n_gpus / (x["ms_per_step"] / 1e3) / 1e9, "ms_per_step": x["ms_per_step"],
                       "e2e_reads_s": n_x * n_gpus / (x["e2e_ms"] / 1e3), "kernel": kernel_name(x["variant"]), "kernel_ms": x["dp_ms"],
                       "ops_per_cell": ops, "pack": pack, "frac": achieved / (peak * pack), "frac_vs_packed_peak": achieved / (peak * 2),
                       "status_ok_reads": x["n_ok"], "gpu_launches": x["launches"], "sub_batches": x["sub_batches"], "pack_retries": x["pack_retries"]}
                if n_gpus == 1:
                    ent["parity"] = oracle_parity(x["c"], x["res"], min(n_chk, n_x), convex=convex)
                    ent["parity"].pop("oracle_seconds", None); ent["parity"].pop("oracle_cells", None)
                extras[key] = ent

    # ---------------- the product's own batch loop (C++ host layer), fed from unpinned memory ----------------
    e2e_api, sharded, e2e_api_packed2 = None, None, None
    if not args.no_api and not args.convex and not args.rustbio:
        try:
            rep = max(1, min(8, 4_000_000 // max(n, 1)))
            br_api, st_api = api_pass(c, [dev], n, passes=3, repeat=rep)
            barrier()
            t = torch.tensor([st_api["seconds"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            api_s = float(t[0])
            same = all(bool((br_api.score_scaled[k * n:(k + 1) * n] == res.score_scaled).all() and (br_api.cigar_len[k * n:(k + 1) * n] == res.cigar_len).all()
                            and (br_api.status[k * n:(k + 1) * n] == res.status).all()) for k in range(rep))
            e2e_api = {"value": n * rep * n_gpus / api_s, "unit": "reads/s", "ms_per_step": 1e3 * api_s / rep, "reads_per_stream": n * rep, "batches": st_api["batches"],
                       "fill_thread_seconds": st_api["fill_seconds"], "sink_thread_seconds": st_api["sink_seconds"],
                       "identical_to_device_resident_results": same,
                       "how": "clique::ShardedAligner::align_reads_span (C++ align_reads loop) on 1 GPU per rank: reads in plain unpinned host memory, "
                              "2 filler threads stage 262144-read batches into page-locked buffers, double-buffered submit / wait, records copied out "
                              "to the caller's arrays; staging copy + H2D + kernels + D2H + copy-out inside the timer; one stream of %d x the step's reads (third pass reported)" % rep}
        except Exception as e:  # noqa: BLE001
            e2e_api = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_packed2 and world == 1 and e2e_api and "error" not in e2e_api:
            # the same loop with AlignerOptions::pack2_upload: the filler threads also pack the staged batch (inside the timer)
            try:
                os.environ["CLQ_SPAN_PACK2"] = "1"
                br_p, st_p = api_pass(c, [dev], n, passes=3, repeat=rep, fillers=4)
                barrier()
                t = torch.tensor([st_p["seconds"]], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                p_s = float(t[0])
                same_p = bool((br_p.score_scaled == br_api.score_scaled).all() and (br_p.cigar_len == br_api.cigar_len).all() and (br_p.status == br_api.status).all()
                              and all(np.array_equal(br_p.cigar(i), br_api.cigar(i)) for i in range(0, n * rep, max(1, n * rep // 4000))))
                e2e_api_packed2 = {"value": n * rep * n_gpus / p_s, "unit": "reads/s", "ms_per_step": 1e3 * p_s / rep, "fill_thread_seconds": st_p["fill_seconds"],
                                   "fillers_per_device": 4, "identical_to_ascii_upload_results": same_p,
                                   "how": "as e2e_api with AlignerOptions::pack2_upload: 4 filler threads stage AND 2-bit pack each batch (ReadBatch::pack2), clq_submit_packed2"}
            except Exception as e:  # noqa: BLE001
                e2e_api_packed2 = {"error": "%s: %s" % (type(e).__name__, e)}
            finally:
                os.environ.pop("CLQ_SPAN_PACK2", None)
        if world > 1:
            # one read stream sharded over all GPUs of the box by the product dispatcher, from ONE process (rank 0); the other
            # ranks wait on the host (gloo), their GPUs idle
            if rank == 0:
                try:
                    ng = torch.cuda.device_count()
                    cs = sorted_c5_stream(args.sharded_reads_per_gpu * ng)
                    br_s, st_s = api_pass(cs, list(range(ng)), len(cs["read_off"]) - 1, passes=2, batch_reads=4096, fillers=2)
                    chk = oracle_parity(cs, br_s, 12)
                    sharded = {"workload": "one C5 stream (mixed 300 bp-5 kb), shortest reads first, %d reads" % (len(cs["read_off"]) - 1), "gpus": ng,
                               "reads_s": st_s["reads"] / st_s["seconds"], "gcups": sum(st_s["device_cells"]) / st_s["seconds"] / 1e9,
                               "seconds": st_s["seconds"], "per_gpu_busy_ms": st_s["device_kernel_ms"], "per_gpu_reads": st_s["device_reads"],
                               "batches": st_s["batches"], "ok": bool(st_s["aligned"] == st_s["reads"] and all(x > 0 for x in st_s["device_reads"]) and chk["mismatches"] == 0),
                               "parity": {"checked_reads": chk["checked_reads"], "mismatches": chk["mismatches"]},
                               "how": "single process, clique::ShardedAligner::align_reads_span over every visible GPU: one cursor, guided batch sizes, no collective"}
                except Exception as e:  # noqa: BLE001
                    sharded = {"error": "%s: %s" % (type(e).__name__, e)}
            dist.barrier(group=cpu_group)

    if rank == 0:
        dp = r["dp_ms"]
        alg_ops, ops = algorithmic_ops(args.workload, args.convex, cells, c, total_bytes)
        achieved = alg_ops / (dp / 1e3) / 1e12
        pack = 2 if (variant & 2) else 1     # s16x2 kernels advance two cells per instruction (peak x pack)
        kernel = kernel_name(variant)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic HBM bytes per launch: raw reads in + results/CIGAR out + 0.5 B/cell of direction bits written once
        tb_cells_hbm = cells if args.workload != "C4" else min(float(sum(len(r) for r in c["refs"])) / len(c["refs"]) * float(total_bytes), cells)
        alg_bytes = float(total_bytes + n * 28 + 4 * int(res.cigar_len.sum()) + (1.0 if args.convex else 0.5) * tb_cells_hbm)
        traffic, ncu = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
            if tr["workload"] == args.workload and (variant & 2):
                traffic = tr["dram_bytes_per_launch"] / tr["reads_per_launch"] * n   # bytes per step, scaled from the ncu capture
                ncu = {k: tr[k] for k in ("alu_pipe_active_pct", "fma_pipe_active_pct", "issue_active_pct", "warps_active_pct", "registers_per_thread", "capture") if k in tr}
        except Exception:
            pass
        line = {
            "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "gcups": gcups,
            "config": {"workload": WORKLOAD_DESC[args.workload].replace("(exhaustive)", "(%s)" % c["search"]) + (" [two-piece affine (convex) gaps, self-pinned]" if args.convex else "") + (" [rust-bio single-reference branch 1/-1/-5/-1 instead of clique's Gotoh, parity unpinned]" if args.rustbio else ""), "reads_per_gpu_per_step": n, "cells_per_gpu_per_step": int(cells),
                       "parallelism": "read-sharded x%d, no collectives" % n_gpus, "status_ok_reads": r["n_ok"], "sub_batches": r["sub_batches"], "pack_retries": r["pack_retries"],
                       "l2": "inputs larger than L2 (%.0f MB of reads + %.0f MB of direction bits per step)" % (total_bytes / 1e6, 0.5 * cells / 1e6)},
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "power_w_max": clk.get("power_w_max"),
                       "samples": clk.get("samples")},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": e2e_ms, "how": "clq_submit/clq_wait on pinned host buffers, %d chunks per step streamed over 2 stream slots (double-buffered across steps)" % r["nch"]},
            "gpu_launches": r["launches"],
            "wall_ms_per_step_device_resident": r["wall_ms"] / args.steps,
            "roofline": {"bound": "int32-alu", "achieved": achieved, "peak": peak * pack, "unit": "TIOP/s", "frac": achieved / (peak * pack),
                         "pack": pack, "peak_int32_alu_pipe": peak, "frac_of_unpacked_int32_peak": achieved / peak, "kernel": kernel,
                         "traffic": traffic, "ops_per_cell": ops, "kernel_ms": dp, "peak_source": peak_how,
                         "hbm": {"achieved": alg_bytes / (dp / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (dp / 1e3) / 1e9 / hbm_peak, "of": "measured" if peaks else "fallback"}},
        }
        if ncu:
            line["roofline"]["ncu"] = ncu
        if e2e_api:
            line["e2e_api"] = e2e_api
        if r.get("packed2"):
            pk = r["packed2"]
            line["e2e_packed2"] = {"value": n * n_gpus / (pk["ms"] / 1e3), "unit": "reads/s", "ms_per_step": pk["ms"], "h2d_bytes_per_step": pk["h2d"],
                                   "d2h_bytes_per_step": pk["d2h"], "exception_bytes": pk["exceptions"], "identical_to_ascii_upload_results": pk["same"],
                                   "host_pack_gb_s_one_thread": pk["host_pack_gb_s_one_thread"],
                                   "how": "as e2e, chunks shipped 2-bit packed (clq_submit_packed2: 0.25 B per base + exception list, expanded on the "
                                          "device by unpack2_kernel); packed ahead of the timer by clq_pack2, whose one-thread rate is reported"}
        if e2e_api_packed2:
            line["e2e_api_packed2"] = e2e_api_packed2
        if sharded:
            line["sharded"] = sharded
        if extras:
            line["configs"] = extras
        if args.convex and n_gpus == 1:
            line["parity"] = oracle_parity(c, res, 64, convex=True)
        elif args.rustbio and n_gpus == 1:
            ns = min(n, 2000)
            t0 = time.perf_counter()
            line["parity"] = oracle_parity(c, res, ns, rustbio=True)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                    "sample": "first %d reads, %.1f s, single-threaded restatement of rust-bio Aligner::global (oracle/, parity unpinned)" % (ns, dt)}
        elif not args.no_cpu_baseline and n_gpus == 1:
            threads = os.cpu_count() or 1
            probe = min(n, 256 if args.workload == "C2" else 16)
            dt, _ = cpu_reference_run(c, probe, threads)
            ns = int(max(probe, min(n, args.cpu_sample_seconds / (dt / probe))))
            par = oracle_parity(c, res, ns, threads=threads)   # the same oracle run is the timed CPU baseline and the parity check
            dt, ocells = par.pop("oracle_seconds"), par.pop("oracle_cells")
            line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": threads, "kind": "port", "gcups": ocells / dt / 1e9,
                                    "sample": "first %d reads of the same workload, %.1f s, f64 port of the reference aligner (oracle/), %d threads" % (ns, dt, threads)}
            line["parity"] = par
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
