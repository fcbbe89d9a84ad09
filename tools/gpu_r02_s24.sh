#!/bin/bash
# round 2, call 24: walker A/B, third pass -- L2 prefetch-size hint (64 / 128 / 256 B) on the direction-bit loads.  LEAN builds.
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s24.txt
cp clique_b200/libclq.so tools/_v/.in_tree.so
trap 'cp tools/_v/.in_tree.so clique_b200/libclq.so' EXIT
b() { timeout -s KILL 200 python bench.py --workload $1 --steps 12 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s %s ms_per_step %.3f  reads/s %.4g  gcups %.1f  e2e %.4g  ok_reads %d  sub_batches %s  pack_retries %s" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"], d["gcups"], d["e2e"]["value"], d["config"]["status_ok_reads"], d["config"].get("sub_batches"), d["config"].get("pack_retries")))' $2 $1 >> $O/r02_s24.txt 2>&1; }
for v in base h64 h128 h256 base h64 h128 h256; do
  cp tools/_v/libclq_$v.so clique_b200/libclq.so
  b C2 $v
done
for v in base h64 h128; do
  cp tools/_v/libclq_$v.so clique_b200/libclq.so
  b C5 $v
  b C3 $v
done
cp tools/_v/libclq_h128.so clique_b200/libclq.so
CLQ_FUZZ_MODES=fixed,fixed,exhaustive,quick CLQ_FUZZ_NO_PACK_P=0.1 timeout -s KILL 60 python tools/fuzz_gpu.py 15 4711 2>&1 | tail -1 >> $O/r02_s24.txt
echo done >> $O/r02_s24.txt
