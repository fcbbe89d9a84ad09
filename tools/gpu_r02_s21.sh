#!/bin/bash
# round 2, call 21: walker A/B -- occupancy (launch bounds 9 / 12 / 16 CTAs per SM) x cp.async look-ahead ring (off / 3 / 7):
# C2 on every variant, C3 + C5 on three of them, focused fuzz on the ring builds.  LEAN builds ((8,40) and (32,32) only).
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s21.txt
cp clique_b200/libclq.so tools/_v/.in_tree.so
trap 'cp tools/_v/.in_tree.so clique_b200/libclq.so' EXIT
b() { timeout -s KILL 200 python bench.py --workload $1 --steps 6 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s %s ms_per_step %.3f  dp_kernel_ms %.3f  walk+rest_ms %.3f  reads/s %.4g  gcups %.1f  e2e %.4g  ok_reads %d  sub_batches %s  pack_retries %s" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["roofline"]["kernel_ms"], d["ms_per_step"] - d["roofline"]["kernel_ms"], d["value"], d["gcups"], d["e2e"]["value"], d["config"]["status_ok_reads"], d["config"].get("sub_batches"), d["config"].get("pack_retries")))' $2 $1 >> $O/r02_s21.txt 2>&1; }
for v in base mb12 mb16 r3mb9 r3mb12 r3mb16 r7mb12 base; do
  cp tools/_v/libclq_$v.so clique_b200/libclq.so
  b C2 $v
done
for v in base r3mb12 r7mb12; do
  cp tools/_v/libclq_$v.so clique_b200/libclq.so
  b C3 $v
  b C5 $v
done
for v in r3mb12 r7mb12; do
  cp tools/_v/libclq_$v.so clique_b200/libclq.so
  echo "fuzz $v" >> $O/r02_s21.txt
  CLQ_FUZZ_MODES=fixed,fixed,exhaustive,quick CLQ_FUZZ_NO_PACK_P=0.1 timeout -s KILL 60 python tools/fuzz_gpu.py 20 4711 2>&1 | tail -1 >> $O/r02_s21.txt
done
cp tools/_v/libclq_r3mb12.so clique_b200/libclq.so
CLQ_FUZZ_WIDE=1 timeout -s KILL 90 python tools/fuzz_gpu.py 40 9921 2>&1 | tail -1 >> $O/r02_s21.txt
echo done >> $O/r02_s21.txt
