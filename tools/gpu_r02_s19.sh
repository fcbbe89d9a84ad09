#!/bin/bash
# round 2, call 19: nibble-packed reference rows in pack_adapt_kernel (two CTAs per SM on 5 kb amplicons): timing + full GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s19.txt
run() { echo "== $*" >> $O/r02_s19.txt; REPRO_DUMP_S=20 timeout -s KILL 35 python tools/repro_c5.py "$@" 2>&1 | grep -v "File\|Thread\|^$" >> $O/r02_s19.txt; }
run 10 len=5000
run 160
if grep -q Timeout $O/r02_s19.txt; then echo "hangs" >> $O/r02_s19.txt; exit 0; fi
for wl in C5 C3; do
  timeout -s KILL 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s ms_per_step %.3f  reads/s %.4g  gcups %.1f  e2e %.4g  dp_kernel_ms %.3f  ok_reads %d  sub_batches %s  pack_retries %s parity %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["gcups"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["status_ok_reads"], d["config"].get("sub_batches"), d["config"].get("pack_retries"), d["config"].get("parity")))' $wl >> $O/r02_s19.txt 2>&1
done
timeout -s KILL 700 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu_r02_s19.log 2>&1; echo "pytest rc=$?" >> $O/r02_s19.txt; tail -5 $O/pytest_gpu_r02_s19.log >> $O/r02_s19.txt
CLQ_FUZZ_WIDE=1 timeout -s KILL 120 python tools/fuzz_gpu.py 80 9918 > $O/fuzz_r02_s19_wide.log 2>&1; tail -1 $O/fuzz_r02_s19_wide.log >> $O/r02_s19.txt
echo done >> $O/r02_s19.txt
