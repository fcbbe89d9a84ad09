#!/bin/bash
# round 2, call 13: geometry sweep for long reads (C3 static PACK, C5 adaptive PACK / int32) on the build with the adaptive kernel at 2 CTAs/SM
cd "$(dirname "$0")/.."
O=gpurun_out
P='import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f variant %s sub_batches %s retries %s" % (d["ms_per_step"], d["value"], d["gcups"], d["roofline"]["kernel"][:40], d["config"].get("sub_batches"), d["config"].get("pack_retries")))'
: > $O/r02_s13.txt
for cfg in 5 4 3 2; do
  echo "== C3 force_cfg=$cfg" >> $O/r02_s13.txt
  CLQ_FORCE_CFG=$cfg timeout 200 python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c "$P" >> $O/r02_s13.txt
done
for cfg in 5 4 3 2; do
  echo "== C5 force_cfg=$cfg" >> $O/r02_s13.txt
  CLQ_FORCE_CFG=$cfg timeout 300 python bench.py --workload C5 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c "$P" >> $O/r02_s13.txt
done
echo done >> $O/r02_s13.txt
