#!/usr/bin/env python3
"""Uploads one C2 batch (default 1 M x 300 bp) in the 2-bit packed form a few times: the program ncu captures unpack2_kernel /
patch2_kernel from (tools/gpu_r02_s31.sh).  Every 500th base is turned into 'N' so that the exception list is exercised.
Also prints the wall time of upload + sync for the ASCII and the packed form (H2D copy included: PCIe-bound, not a kernel time)."""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from clique_b200 import Aligner, Reference, ReferenceManager, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
c = synth.config_c2(n)
total = int(c["read_off"][-1])
al = Aligner(device=0, max_reads=n, max_read_bytes=total + 64, max_read_len=1 << 15, cigar_ops_per_read=12, n_slots=1)
al.set_references(ReferenceManager([Reference(r, nm) for r, nm in zip(c["refs"], c["ref_names"])]))
hb = al.alloc_pinned(total, np.uint8)
hb[:] = c["read_bytes"][:total]
hb[::500] = ord("N")
words = al.alloc_pinned((total + 15) // 16 + 4, np.uint32)
t0 = time.perf_counter()
pk = al.pack_reads(hb, total, out_words=words)
t_pack = time.perf_counter() - t0
for form, src in (("ascii", hb), ("packed2", pk)):
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        al.upload(0, src, c["read_off"], c["fixed_ref"])
        al.sync(0)
        ts.append(time.perf_counter() - t0)
    nbytes = total if form == "ascii" else 4 * ((total + 15) // 16) + 9 * len(pk.exc_pos)
    print("%s upload + sync: %.3f ms (best of 4), read bytes on the wire %d" % (form, 1e3 * min(ts), nbytes))
print("bases %d, exceptions %d, host pack %.2f GB/s (one thread)" % (total, len(pk.exc_pos), total / t_pack / 1e9))
al.close()
