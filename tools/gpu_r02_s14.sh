#!/bin/bash
# round 2, call 14: long reads on (8,40) with column stripes (static + adaptive s16x2), narrow last stripe there: full suite, fuzz, canary, C3 / C5
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s14.log 2>&1; echo "pytest rc=$?" > $O/r02_s14.txt
CLQ_FUZZ_VERBOSE=1 CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 90 4242 > $O/fuzz_r02_s14_wide.log 2>&1; echo "fuzz wide rc=$?" >> $O/r02_s14.txt
timeout 200 python tools/fuzz_gpu.py 60 1234 > $O/fuzz_r02_s14.log 2>&1; echo "fuzz rc=$?" >> $O/r02_s14.txt
AB_WORKLOADS="C2 C3 C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r14 > $O/ab_r02_s14.txt 2>&1
bash tools/canary_gpu.sh > $O/canary_r02_s14.log 2>&1; echo "canary rc=$?" >> $O/r02_s14.txt
echo done >> $O/r02_s14.txt
