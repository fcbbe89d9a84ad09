#!/bin/bash
# round 2, call 26: (1) the share of the traceback walk per config (debug_flags = 1 skips the walk kernel: step time with / without),
# (2) the window canary on a HEAD build (long reads on the static (8,40) window with column stripes, overlapped sub-batches),
# (3) a longer fuzz campaign on the shipped library.
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s26.txt
b() { env "${@:2}" timeout -s KILL 200 python bench.py --workload $1 --steps 8 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s [%s] ms_per_step %.3f  reads/s %.4g  gcups %.1f  sub_batches %s" % (sys.argv[1], " ".join(sys.argv[2:]), d["ms_per_step"], d["value"], d["gcups"], d["config"].get("sub_batches")))' $1 "${@:2}" >> $O/r02_s26.txt 2>&1; }
for wl in C2 C3 C5; do
  b $wl CLQ_X=0
  b $wl CLQ_DEBUG_FLAGS=1
done
bash tools/canary_gpu.sh > $O/canary_r02_s26.log 2>&1; echo "canary rc=$?" >> $O/r02_s26.txt; tail -1 $O/canary_r02_s26.log >> $O/r02_s26.txt
timeout -s KILL 200 python tools/fuzz_gpu.py 150 20263 > $O/fuzz_r02_final_seed20263.log 2>&1; tail -1 $O/fuzz_r02_final_seed20263.log >> $O/r02_s26.txt
echo done >> $O/r02_s26.txt
