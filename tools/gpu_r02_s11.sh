#!/bin/bash
# round 2, call 11: adaptive kernel without the spurious retries of padded pairs (3 vs 2 CTAs/SM), the complete bench line, canary
cd "$(dirname "$0")/.."
O=gpurun_out
AB_WORKLOADS="C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r11 a2 > $O/ab_r02_s11.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s11.log 2>&1; echo "pytest rc=$?" > $O/r02_s11.txt
(time timeout 900 python bench.py > $O/bench_r02_s11_full.json 2> $O/bench_r02_s11_full.err) 2>> $O/r02_s11.txt; echo "bench rc=$?" >> $O/r02_s11.txt
bash tools/canary_gpu.sh > $O/canary_r02_s11.log 2>&1; echo "canary rc=$?" >> $O/r02_s11.txt
echo done >> $O/r02_s11.txt
