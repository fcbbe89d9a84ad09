#!/usr/bin/env python3
"""Markdown table of one bench.py JSON line (headline + the `configs` block), for DESIGN.md section 5.
usage: tools/results_table.py profiles/bench_r02_final.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).readline())
rows = [("C2 " + ("(%d GPUs)" % d["n_gpus"] if d["n_gpus"] > 1 else "(headline)"), d["config"]["reads_per_gpu_per_step"], d["value"], d["gcups"], d["e2e"]["value"],
         d["roofline"]["kernel"], d["roofline"]["frac"], d.get("parity", {}).get("mismatches"), d.get("parity", {}).get("checked_reads"), d["config"].get("pack_retries"))]
for k, x in d.get("configs", {}).items():
    if "error" in x:
        rows.append((k, "-", "error: " + x["error"], "", "", "", "", "", "", ""))
        continue
    rows.append((k, x["reads_per_gpu_per_step"], x["reads_s"], x["gcups"], x["e2e_reads_s"], x["kernel"], x["frac_vs_packed_peak"],
                 x.get("parity", {}).get("mismatches"), x.get("parity", {}).get("checked_reads"), x.get("pack_retries")))
print("| config | reads / step | reads/s (device) | GCUPS | reads/s (e2e, pinned host buffers) | kernel | ops-fraction of the packed INT32 peak | oracle spot check |")
print("|---|---|---|---|---|---|---|---|")
for r in rows:
    if isinstance(r[2], str):
        print("| %s | %s | %s | | | | | |" % (r[0], r[1], r[2]))
        continue
    print("| %s | %s | %.4g | %.0f | %.4g | %s | %.2f | %s / %s |" % (r[0], r[1], r[2], r[3], r[4], r[5].split(" (")[0], r[6], r[7], r[8]))
for k in ("e2e_api", "sharded", "cpu_baseline"):
    if k in d:
        x = dict(d[k]); x.pop("how", None); x.pop("sample", None)
        print("\n`%s`: `%s`" % (k, json.dumps(x)))
print("\nclocks: `%s`; `roofline.ncu`: `%s`" % (json.dumps(d["clocks"]), json.dumps({k: v for k, v in d["roofline"].get("ncu", {}).items() if k != "capture"})))
