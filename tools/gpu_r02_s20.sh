#!/bin/bash
# round 2, call 20: overlapped sub-batches (two streams, contiguous longest-first pieces): C5 with / without, parity, fuzz
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s20.txt
run() { echo "== $*" >> $O/r02_s20.txt; REPRO_DUMP_S=20 timeout -s KILL 35 python tools/repro_c5.py "$@" 2>&1 | grep -v "File\|Thread\|^$" >> $O/r02_s20.txt; }
run 160
run 160 max_scratch_bytes=16777216
if grep -q Timeout $O/r02_s20.txt; then echo "hangs" >> $O/r02_s20.txt; exit 0; fi
b() { echo "-- $*" >> $O/r02_s20.txt; env "$@" timeout -s KILL 200 python bench.py --workload C5 --steps 5 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("C5 ms_per_step %.3f  reads/s %.4g  gcups %.1f  e2e %.4g  dp_kernel_ms %.3f  ok_reads %d  sub_batches %s  pack_retries %s parity %s" % (d["ms_per_step"], d["value"], d["gcups"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["status_ok_reads"], d["config"].get("sub_batches"), d["config"].get("pack_retries"), d["config"].get("parity")))' >> $O/r02_s20.txt 2>&1; }
b CLQ_X=0
b CLQ_NO_OVERLAP=1
b CLQ_MAX_SCRATCH_BYTES=51539607552
b CLQ_NO_LONG8=1
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 -k "adaptive or sub or grouped or c5 or window" > $O/pytest_gpu_r02_s20.log 2>&1; echo "pytest rc=$?" >> $O/r02_s20.txt; tail -3 $O/pytest_gpu_r02_s20.log >> $O/r02_s20.txt
CLQ_FUZZ_WIDE=1 timeout -s KILL 150 python tools/fuzz_gpu.py 110 9919 > $O/fuzz_r02_s20_wide.log 2>&1; tail -1 $O/fuzz_r02_s20_wide.log >> $O/r02_s20.txt
echo done >> $O/r02_s20.txt
