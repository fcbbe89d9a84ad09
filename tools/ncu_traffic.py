#!/usr/bin/env python3
"""profiles/traffic_rNN.json from an `ncu --set full` capture of the C2 fill kernel (read here, no GPU needed): per-launch DRAM bytes
and the pipe / issue utilisation bench.py copies into roofline.traffic and roofline.ncu.
usage: tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/traffic_r02.json <reads per launch> "<capture note>" """
import csv
import io
import json
import subprocess
import sys

rep, out, reads, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
fills = [dict(zip(hdr, r)) for r in rows[2:] if "pack_kernel" in dict(zip(hdr, r)).get("Kernel Name", "")]
walks = [dict(zip(hdr, r)) for r in rows[2:] if "walk_kernel" in dict(zip(hdr, r)).get("Kernel Name", "")]
f = lambda d, k: float(d[k].replace(",", ""))
units = dict(zip(hdr, rows[1]))
scale = lambda k: {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[k]]
d = fills[-1]
j = {"workload": "C2", "kernel": d["Kernel Name"].split("(")[0], "reads_per_launch": reads, "capture": note,
     "dram_bytes_per_launch": f(d, "dram__bytes_read.sum") * scale("dram__bytes_read.sum") + f(d, "dram__bytes_write.sum") * scale("dram__bytes_write.sum"),
     "kernel_ms_under_ncu": f(d, "gpu__time_duration.sum"),
     "alu_pipe_active_pct": f(d, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
     "fma_pipe_active_pct": f(d, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
     "issue_active_pct": f(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "warps_active_pct": f(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
     "registers_per_thread": int(f(d, "launch__registers_per_thread"))}
if walks:
    w = walks[-1]
    j["walk"] = {"kernel_ms_under_ncu": f(w, "gpu__time_duration.sum"),
                 "dram_bytes_read_per_read": f(w, "dram__bytes_read.sum") * scale("dram__bytes_read.sum") / reads,
                 "long_scoreboard_stall_per_issue": f(w, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")}
json.dump(j, open(out, "w"), indent=1)
print(json.dumps(j, indent=1))
