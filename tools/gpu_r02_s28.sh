#!/bin/bash
# round 2, call 28 (last budget): HEAD confirmation -- the full GPU suite, the complete bench line, the share of the traceback
# walk per config (debug_flags = 1 skips the walk kernel: step time with / without), one fuzz sweep.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s28.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $O/r02_s28.txt
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu_r02_head.log 2>&1; echo "pytest rc=$?" >> $O/r02_s28.txt; tail -3 $O/pytest_gpu_r02_head.log >> $O/r02_s28.txt
(time timeout -s KILL 300 python bench.py > $O/bench_r02_head.json 2> $O/bench_r02_head.err) 2>> $O/r02_s28.txt; echo "bench rc=$?" >> $O/r02_s28.txt
b() { env "${@:2}" timeout -s KILL 120 python bench.py --workload $1 --steps 8 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s [%s] ms_per_step %.3f  reads/s %.4g  gcups %.1f  sub_batches %s" % (sys.argv[1], " ".join(sys.argv[2:]), d["ms_per_step"], d["value"], d["gcups"], d["config"].get("sub_batches")))' $1 "${@:2}" >> $O/r02_s28.txt 2>&1; }
for wl in C2 C3 C5; do
  b $wl CLQ_X=0
  b $wl CLQ_DEBUG_FLAGS=1
done
timeout -s KILL 90 python tools/fuzz_gpu.py 70 20263 > $O/fuzz_r02_head_seed20263.log 2>&1; tail -1 $O/fuzz_r02_head_seed20263.log >> $O/r02_s28.txt
echo done >> $O/r02_s28.txt
