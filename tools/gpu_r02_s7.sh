#!/bin/bash
# round 2, call 7: lean general path (predicated boundary column, nested blocks) on the PACK and int32 kernels, warp-per-pair walker for
# long reads, window-edge tests: full GPU suite, fuzz, A/B on C2 / C3 / C5 against the previous build
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s7.log 2>&1; echo "pytest rc=$?" > $O/r02_s7.txt
timeout 200 python tools/fuzz_gpu.py 60 1234 > $O/fuzz_r02_s7.log 2>&1; echo "fuzz rc=$?" >> $O/r02_s7.txt
CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 60 4242 > $O/fuzz_r02_s7_wide.log 2>&1; echo "fuzz wide rc=$?" >> $O/r02_s7.txt
AB_WORKLOADS="C2 C3 C5" AB_STEPS=4 FUZZ_SECONDS=5 timeout 1200 tools/ab_variants.sh run pf12 r6 > $O/ab_r02_s7.txt 2>&1
echo done >> $O/r02_s7.txt
