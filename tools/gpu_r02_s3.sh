#!/bin/bash
# round 2, call 3: full GPU suite (incl. the span dispatcher tests), A/B of the long-read path, the complete bench line
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s3.log 2>&1; echo "pytest rc=$?" > $O/r02_s3.txt
AB_WORKLOADS="C3 C5" AB_STEPS=3 FUZZ_SECONDS=5 timeout 900 tools/ab_variants.sh run base r2full > $O/ab_r02_s3_long.txt 2>&1
(time timeout 900 python bench.py > $O/bench_r02_s3_full.json 2> $O/bench_r02_s3_full.err) 2>> $O/r02_s3.txt; echo "bench rc=$?" >> $O/r02_s3.txt
