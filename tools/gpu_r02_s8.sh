#!/bin/bash
# round 2, call 8: fixes after call 7 (padding positions in the convex kernel, thread-per-pair walker again, sub-batches dealt
# round-robin): full GPU suite, fuzz, canary with diagnostics, C3 / C5 with and without the adaptive kernel, C5 launch list
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s8.log 2>&1; echo "pytest rc=$?" > $O/r02_s8.txt
timeout 200 python tools/fuzz_gpu.py 60 1234 > $O/fuzz_r02_s8.log 2>&1; echo "fuzz rc=$?" >> $O/r02_s8.txt
CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 60 4242 > $O/fuzz_r02_s8_wide.log 2>&1; echo "fuzz wide rc=$?" >> $O/r02_s8.txt
AB_WORKLOADS="C2 C3 C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r8 > $O/ab_r02_s8.txt 2>&1
for v in "CLQ_NO_ADAPT=1" "CLQ_NO_ADAPT=0"; do
  echo "== C5 $v" >> $O/r02_s8.txt
  env $v timeout 200 python bench.py --workload C5 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f sub_batches %s retries %s" % (d["ms_per_step"], d["value"], d["gcups"], d["config"].get("sub_batches"), d["config"].get("pack_retries")))' >> $O/r02_s8.txt
done
CMD="python bench.py --workload C5 --reads 12000 --steps 2 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout 200 $CMD > $O/plain_C5.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r02_s8_C5.csv $CMD > $O/ncu_l_C5.log 2>&1
bash tools/canary_gpu.sh > $O/canary_r02_s8.log 2>&1; echo "canary rc=$?" >> $O/r02_s8.txt
echo done >> $O/r02_s8.txt
