#!/bin/bash
# round 2, call 12: adaptive kernel 3 vs 2 CTAs/SM (fresh builds), its ncu capture + launch list on C5, canary with real-cell masking
cd "$(dirname "$0")/.."
O=gpurun_out
AB_WORKLOADS="C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r12 a2 > $O/ab_r02_s12.txt 2>&1
bash tools/canary_gpu.sh > $O/canary_r02_s12.log 2>&1; echo "canary rc=$?" > $O/r02_s12.txt
CMD="python bench.py --workload C5 --reads 12000 --steps 2 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout 200 $CMD > $O/plain_C5.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r02_s12_C5.csv $CMD > $O/ncu_l_C5.log 2>&1
CMD2="python bench.py --workload C5 --reads 12000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout 200 $CMD2 > $O/plain2_C5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pack_adapt_kernel' -c 1 -f -o $O/prof_r02_s12_C5 $CMD2 > $O/ncu_f_C5.log 2>&1
echo done >> $O/r02_s12.txt
