#!/bin/bash
# round 2, call 5: MADD (static row slope) + pinned nibble chain: full GPU suite, fuzz, A/B on C2 (no_madd toggle), C3 / C5 with
# the time-transposed bit layout on the long-read geometries
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s5.log 2>&1; echo "pytest rc=$?" > $O/r02_s5.txt
timeout 200 python tools/fuzz_gpu.py 60 9911 > $O/fuzz_r02_s5.log 2>&1; echo "fuzz rc=$?" >> $O/r02_s5.txt
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-live-peak --no-extra --no-api"
for v in "" "CLQ_NO_MADD=1"; do
  echo "== C2 $v" >> $O/r02_s5.txt
  env $v timeout 120 $B 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f e2e %.4g ok %d" % (d["ms_per_step"], d["value"], d["gcups"], d["e2e"]["value"], d["config"]["status_ok_reads"]))' >> $O/r02_s5.txt
done
for wl in "C4 --search exhaustive --reads 50000" "C4 --search quick --reads 400000"; do
  echo "== $wl" >> $O/r02_s5.txt
  timeout 120 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f" % (d["ms_per_step"], d["value"], d["gcups"]))' >> $O/r02_s5.txt
done
AB_WORKLOADS="C3 C5" AB_STEPS=3 FUZZ_SECONDS=15 timeout 900 tools/ab_variants.sh run r3 r4 > $O/ab_r02_s5_long.txt 2>&1
echo done >> $O/r02_s5.txt
