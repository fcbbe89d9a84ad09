#!/bin/bash
# round 2, call 10: score stage keeps the upload order (wide-fuzz crash), 128-bit boundary-column rows; why does the adaptive kernel retry?
cd "$(dirname "$0")/.."
O=gpurun_out
CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 90 4242 > $O/fuzz_r02_s10_wide.log 2>&1; echo "fuzz wide rc=$?" > $O/r02_s10.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s10.log 2>&1; echo "pytest rc=$?" >> $O/r02_s10.txt
AB_WORKLOADS="C3 C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r10 > $O/ab_r02_s10.txt 2>&1
# diagnostics build: which check makes the adaptive kernel hand a pair to the retry pass?
cp clique_b200/libclq.so tools/_v/.in_tree.so; cp tools/_v/libclq_canary.so clique_b200/libclq.so
timeout 300 python bench.py --workload C5 --reads 30000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api > $O/adapt_diag.json 2> $O/adapt_diag.err
grep -h "adapt retry" $O/adapt_diag.json $O/adapt_diag.err | sort | uniq -c | sort -rn | head -40 > $O/adapt_retry_reasons.txt
cp tools/_v/.in_tree.so clique_b200/libclq.so
CMD="python bench.py --workload C3 --reads 40000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout 200 $CMD > $O/plain_C3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pack_kernel' -c 2 -f -o $O/prof_r02_s10_C3 $CMD > $O/ncu_f_C3.log 2>&1
echo done >> $O/r02_s10.txt
