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
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from clique_b200 import AffineScoring, Aligner, Reference, ReferenceManager, RustBioScoring, TwoPieceScoring

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    dev = local_rank if world > 1 else 0
    n = args.reads or default_reads(args.workload)
    c = make_workload(args.workload, n, args.search)
    # weak scaling: every rank aligns its own n reads per step (same generator, rank-specific rotation of the batch)
    if rank:
        k = (rank * 7919) % n
        lens = (c["read_off"][1:] - c["read_off"][:-1]).astype(np.int64)
        order = np.roll(np.arange(n), k)
        if (lens == lens[0]).all():
            L = int(lens[0])
            c["read_bytes"] = np.roll(c["read_bytes"].reshape(n, L), k, axis=0).ravel()
            if c["fixed_ref"] is not None:
                c["fixed_ref"] = np.roll(c["fixed_ref"], k)
        del order
    total_bytes = int(c["read_off"][-1])
    ops_per_read = {"C2": 12, "C3": 400, "C4": 12, "C5": 700}[args.workload]
    al = Aligner(device=dev, max_reads=n, max_read_bytes=total_bytes + 64, max_read_len=1 << 15, max_refs=max(64, len(c["refs"])),
                 cigar_ops_per_read=ops_per_read, n_slots=2)
    al.set_references(ReferenceManager([Reference(r, nm) for r, nm in zip(c["refs"], c["ref_names"])]))
    sc = TwoPieceScoring(10, -9, 9, -20, -2, -40, -1) if args.convex else AffineScoring(*c["scoring"])
    sci = sc.to_int()
    if args.rustbio:
        assert c["search"] == "fixed" and not args.convex, "--rustbio is the single-reference branch"
        sci = RustBioScoring()   # Aligner.launch adds CLQ_RUSTBIO for this scoring type
        c["band"] = "maxlen"
    for opt in ("force_cfg", "force_generic", "debug_flags"):     # experiment knobs, e.g. CLQ_FORCE_CFG=3
        if os.environ.get("CLQ_" + opt.upper()):
            al.set_option(opt, int(os.environ["CLQ_" + opt.upper()]))
    score_only = bool(int(os.environ.get("CLQ_SCORE_ONLY", "0")))

    # pinned host buffers for the e2e path, split into chunks that alternate over the two stream slots
    h_bytes = al.alloc_pinned(total_bytes, np.uint8)
    h_bytes[:] = c["read_bytes"][:total_bytes]
    nch = max(1, args.e2e_chunks)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput: inputs in HBM, launches only ----------------
    al.upload(0, h_bytes, c["read_off"], c["fixed_ref"])
    al.sync(0)
    for _ in range(args.warmup):
        al.launch(0, sci, c["search"], c["band"], score_only)
        al.sync(0)
    clocks = ClockSampler(dev)
    clocks.start()
    barrier()
    step_ms, dp_ms, launches, cells, variant = [], [], 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        al.launch(0, sci, c["search"], c["band"], score_only)
        st = al.stats(0)       # synchronises the slot's stream; times come from CUDA events on that stream
        step_ms.append(st["kernel_ms"]); dp_ms.append(st["dp_ms"]); launches += st["launches"]; cells = st["cells"]; variant = st["variant"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk = clocks.stop()
    res = al.wait(0, copy=True)
    n_ok = int((res.status == 0).sum())
    dev_ms_total = float(np.sum(step_ms))
    t = torch.tensor([dev_ms_total, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_total, wall_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms_total / args.steps
    value = n * n_gpus / (ms_per_step / 1e3)
    gcups = cells * n_gpus / (ms_per_step / 1e3) / 1e9

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
    for _ in range(args.warmup):
        e2e_step()

    def e2e_run(steps):
        """K steps as the product runs them: one double-buffered stream of chunks (chunk k+1 uploads while chunk k computes,
        also across step boundaries); every chunk's inputs are copied from pinned host memory and its results read back."""
        busy, k, last = [False, False], 0, None
        for _step in range(steps):
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
    e2e_run(args.steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0]) / args.steps
    e2e_value = n * n_gpus / (e2e_ms / 1e3)

    if rank == 0:
        peak, peak_how = int32_peak(not args.no_live_peak)
        dp = float(np.mean(dp_ms))
        ops = (30 if args.convex else OPS_PER_CELL_TB) if args.workload != "C4" else None   # two-piece: 20 / 30
        if args.workload == "C4":   # n score-only fills + 1 traceback fill per read (cells counts both)
            tb_cells = float(sum(len(r) for r in c["refs"])) / len(c["refs"]) * float(total_bytes)
            alg_ops = cells * OPS_PER_CELL_SCORE + min(tb_cells, cells) * (OPS_PER_CELL_TB - OPS_PER_CELL_SCORE)
        else:
            alg_ops = cells * ops
        achieved = alg_ops / (dp / 1e3) / 1e12
        pack = 2 if (variant & 2) else 1     # s16x2 kernels advance two cells per instruction (peak x pack)
        kernel = {0: "generic int32", 1: "FAST int32 (PRMT profile + DPX)", 3: "PACK s16x2 (two reads per lane group, DPX)",
                  5: "CONVEX int32 (two-piece affine, DPX)", 7: "CONVEX PACK s16x2 (two-piece affine, two reads per lane group, DPX)"}.get(variant & 7, "variant %d" % (variant & 15))
        if variant & 16:
            kernel += " [rust-bio global semantics]"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        lens = (c["read_off"][1:] - c["read_off"][:-1]).astype(np.int64)
        # algorithmic HBM bytes per launch: raw reads in + results/CIGAR out + 0.5 B/cell of direction bits written once
        tb_cells_hbm = cells if args.workload != "C4" else min(float(sum(len(r) for r in c["refs"])) / len(c["refs"]) * float(total_bytes), cells)
        alg_bytes = float(total_bytes + n * 28 + 4 * int(res.cigar_len.sum()) + (1.0 if args.convex else 0.5) * tb_cells_hbm)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
            if tr["workload"] == args.workload and (variant & 2):
                traffic = tr["dram_bytes_per_launch"] / tr["reads_per_launch"] * n   # bytes per step, scaled from the ncu capture
        except Exception:
            pass
        line = {
            "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "gcups": gcups,
            "config": {"workload": WORKLOAD_DESC[args.workload].replace("(exhaustive)", "(%s)" % c["search"]) + (" [two-piece affine (convex) gaps, self-pinned]" if args.convex else "") + (" [rust-bio single-reference branch 1/-1/-5/-1 instead of clique's Gotoh, parity unpinned]" if args.rustbio else ""), "reads_per_gpu_per_step": n, "cells_per_gpu_per_step": int(cells),
                       "parallelism": "read-sharded x%d, no collectives" % n_gpus, "status_ok_reads": n_ok,
                       "l2": "inputs larger than L2 (%.0f MB of reads + %.0f MB of direction bits per step)" % (total_bytes / 1e6, 0.5 * cells / 1e6)},
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"], "power_w_max": clk.get("power_w_max"),
                       "samples": clk.get("samples")},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "how": "clq_submit/clq_wait on pinned host buffers, %d chunks per step streamed over 2 stream slots (double-buffered across steps)" % nch},
            "gpu_launches": int(launches),
            "wall_ms_per_step_device_resident": wall_ms / args.steps,
            "roofline": {"bound": "int32-alu", "achieved": achieved, "peak": peak * pack, "unit": "TIOP/s", "frac": achieved / (peak * pack),
                         "pack": pack, "peak_int32_alu_pipe": peak, "frac_of_unpacked_int32_peak": achieved / peak, "kernel": kernel,
                         "traffic": traffic, "ops_per_cell": ops if ops else "12 score-only + 18 traceback", "kernel_ms": dp, "peak_source": peak_how,
                         "hbm": {"achieved": alg_bytes / (dp / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (dp / 1e3) / 1e9 / hbm_peak, "of": "measured" if peaks else "fallback"}},
        }
        if args.convex and n_gpus == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import _oracle as O
            ocv = O.Convex(10, -9, 9, -20, -2, -40, -1, -100000)
            bad, ns = 0, 64
            for i in range(ns):
                rd = bytes(c["read_bytes"][int(c["read_off"][i]):int(c["read_off"][i + 1])])
                w = O.convex_align_pair(c["refs"][int(c["fixed_ref"][i]) if c["fixed_ref"] is not None else int(res.ref_index[i])], rd, ocv)
                if w["score"] != int(res.score_scaled[i]) or O.cigar_str(w["cigar"]) != res.cigar_string(i):
                    bad += 1
            line["parity"] = {"checked_reads": ns, "mismatches": bad, "oracle": "orc_convex_align_pair (self-pinned)"}
        elif args.rustbio and n_gpus == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import _oracle as O
            bad, ns = 0, min(n, 2000)
            t0 = time.perf_counter()
            for i in range(ns):
                rd = bytes(c["read_bytes"][int(c["read_off"][i]):int(c["read_off"][i + 1])])
                w = O.rustbio_global(c["refs"][int(c["fixed_ref"][i])], rd)
                if w["score"] != int(res.score_scaled[i]) or O.cigar_str(w["cigar"]) != res.cigar_string(i):
                    bad += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                    "sample": "first %d reads, %.1f s, single-threaded restatement of rust-bio Aligner::global (oracle/, parity unpinned)" % (ns, dt)}
            line["parity"] = {"checked_reads": ns, "mismatches": bad, "oracle": "orc_rustbio_global (parity unpinned)"}
        elif not args.no_cpu_baseline and n_gpus == 1:
            threads = os.cpu_count() or 1
            probe = min(n, 256 if args.workload == "C2" else 16)
            dt, _ = cpu_reference_run(c, probe, threads)
            ns = int(max(probe, min(n, args.cpu_sample_seconds / (dt / probe))))
            dt, out = cpu_reference_run(c, ns, threads)
            line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": threads, "kind": "port", "gcups": out["cells"] / dt / 1e9,
                                    "sample": "first %d reads of the same workload, %.1f s, f64 port of the reference aligner (oracle/), %d threads" % (ns, dt, threads)}
            # parity spot-check of the timed batch against the same oracle run
            bad = 0
            for i in range(ns):
                o, l = int(out["cigar_off"][i]), int(out["cigar_len"][i])
                if int(res.score_scaled[i]) != out["score"][i] * res.scale or not np.array_equal(res.cigar(i), out["cigar_pool"][o:o + l]) \
                        or int(res.ref_index[i]) != int(out["ref_index"][i]):
                    bad += 1
            line["parity"] = {"checked_reads": ns, "mismatches": bad}
        print(json.dumps(line))
    al.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
