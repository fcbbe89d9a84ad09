"""Host-side timeline of the e2e loop (experiments): per call wall times of submit / wait for 1..4 chunks per step."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clique_b200 import AffineScoring, Aligner, Reference, ReferenceManager, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = synth.config_c2(n)
total = int(c["read_off"][-1])
al = Aligner(device=0, max_reads=n, max_read_bytes=total + 64, max_read_len=1 << 15, cigar_ops_per_read=12, n_slots=2)
al.set_references(ReferenceManager([Reference(r, nm) for r, nm in zip(c["refs"], c["ref_names"])]))
sci = AffineScoring(*c["scoring"]).to_int()
h = al.alloc_pinned(total, np.uint8)
h[:] = c["read_bytes"][:total]
bounds = [n * i // nch for i in range(nch + 1)]
chunks = []
for i in range(nch):
    lo, hi = bounds[i], bounds[i + 1]
    off = al.alloc_pinned(hi - lo + 1, np.uint64); off[:] = c["read_off"][lo:hi + 1] - c["read_off"][lo]
    fr = al.alloc_pinned(hi - lo, np.int32); fr[:] = 0
    chunks.append((h[int(c["read_off"][lo]):int(c["read_off"][hi])], off, fr))
for _ in range(2):
    for i, (rb, off, fr) in enumerate(chunks):
        al.submit(i % 2, rb, off, sci, "fixed", "readlen", fixed_ref=fr); al.wait(i % 2, copy=False)
busy, k = [False, False], 0
t00 = time.perf_counter()
log = []
for step in range(6):
    for rb, off, fr in chunks:
        s0 = k % 2
        if busy[s0]:
            t0 = time.perf_counter(); al.wait(s0, copy=False); log.append(("wait", s0, 1e3 * (time.perf_counter() - t0)))
        t0 = time.perf_counter(); al.upload(s0, rb, off, fr); t1 = time.perf_counter(); al.launch(s0, sci, "fixed", "readlen"); t2 = time.perf_counter()
        al._check(al.lib.clq_download(al.ctx, s0)); t3 = time.perf_counter()
        log.append(("upload", s0, 1e3 * (t1 - t0))); log.append(("launch", s0, 1e3 * (t2 - t1))); log.append(("download", s0, 1e3 * (t3 - t2)))
        busy[s0] = True; k += 1
for s0 in (k % 2, (k + 1) % 2):
    if busy[s0]:
        t0 = time.perf_counter(); al.wait(s0, copy=False); log.append(("wait", s0, 1e3 * (time.perf_counter() - t0)))
tot = 1e3 * (time.perf_counter() - t00)
print("total ms", tot, "per step", tot / 6)
for e in log:
    print("%-9s slot %d %8.3f ms" % e)
