#!/usr/bin/env python3
"""A/B of the span dispatcher (ShardedAligner::align_reads_span) on one C5 stream sorted shortest-first: claim order (front /
longest first / two-ended), batch size, chained vs independent stream slots; every configuration's records are compared with the
first one's.  Needs >= 1 GPU; run on a 2-GPU box:  python tools/span_experiment.py [reads_per_gpu]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

ng = torch.cuda.device_count()
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
cs = bench.sorted_c5_stream(per_gpu * max(ng, 1))
n = len(cs["read_off"]) - 1
cells = float(sum(len(cs["refs"][r]) * int(l) for r, l in zip(cs["fixed_ref"], (cs["read_off"][1:] - cs["read_off"][:-1]))))
print("stream: %d reads, %.1f MB, %.1f Gcells, %d GPUs" % (n, cs["read_off"][-1] / 1e6, cells / 1e9, ng), flush=True)
base = None


def run(devs, order, batch_reads, serialize, label=""):
    global base
    for k in ("CLQ_SPAN_ORDER", "CLQ_SPAN_SERIALIZE_SLOTS", "CLQ_SPAN_MAX_SCRATCH_BYTES"):
        os.environ.pop(k, None)
    os.environ["CLQ_SPAN_ORDER"] = order
    if not serialize:
        os.environ["CLQ_SPAN_SERIALIZE_SLOTS"] = "0"
        os.environ["CLQ_SPAN_MAX_SCRATCH_BYTES"] = str(40 << 30)
    t0 = time.time()
    try:
        br, st = bench.api_pass(cs, devs, n, passes=2, batch_reads=batch_reads, fillers=2)
    except Exception as e:  # noqa: BLE001
        print("%-8s devs %s order %-7s batch_reads %6d serialize %d: ERROR %s" % (label, devs, order, batch_reads, serialize, e), flush=True)
        return
    same = None
    if base is None:
        base = br
    else:
        same = bool((br.score_scaled == base.score_scaled).all() and (br.cigar_len == base.cigar_len).all() and (br.status == base.status).all()
                    and (br.ref_index == base.ref_index).all() and all(np.array_equal(br.cigar(i), base.cigar(i)) for i in range(0, n, 997)))
    print("%-8s devs %s order %-7s batch_reads %6d serialize %d: %.1f ms  %.0f GCUPS  batches %d  busy_ms %s  reads %s  same_as_first %s  (call %.1f s)" %
          (label, devs, order, batch_reads, serialize, 1e3 * st["seconds"], cells / st["seconds"] / 1e9, st["batches"],
           ["%.0f" % x for x in st["device_kernel_ms"]], st["device_reads"], same, time.time() - t0), flush=True)


one = [0]
run(one, "front", 4096, 1, "1gpu")
run(one, "front", 65536, 1, "1gpu")
run(one, "two", 65536, 1, "1gpu")
if ng > 1:
    alld = list(range(ng))
    for order in ("front", "longest", "two"):
        for brd in (4096, 32768):
            for ser in (1, 0):
                run(alld, order, brd, ser)
