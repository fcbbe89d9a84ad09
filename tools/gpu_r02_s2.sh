#!/bin/bash
# round 2, call 2: full GPU suite on the restructured pack kernel, A/B against the round-1 build, first multi-config bench line
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s2.log 2>&1; echo "pytest rc=$?" > $O/r02_s2.txt
AB_WORKLOADS="C2 C3" FUZZ_SECONDS=20 timeout 900 tools/ab_variants.sh run base r1 > $O/ab_r02_s2_restructure.txt 2>&1
(time timeout 600 python bench.py > $O/bench_r02_s2_full.json 2> $O/bench_r02_s2_full.err) 2>> $O/r02_s2.txt; echo "bench rc=$?" >> $O/r02_s2.txt
