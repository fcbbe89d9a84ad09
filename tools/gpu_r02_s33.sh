#!/bin/bash
# round 2, call 33 (the last GPU minutes): ncu --set full of the C4 exhaustive score stage (pack_kernel<.., TB = 0>: score-only s16x2 fill)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s33.txt
CMD="python bench.py --workload C4 --reads 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api --no-packed2"
timeout -s KILL 100 $CMD > $O/plain_c4.log 2>&1; echo "plain rc=$?" >> $O/r02_s33.txt
timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k regex:'pack_kernel' -c 3 -f -o $O/prof_r02_c4_score $CMD > $O/ncu_c4.log 2>&1; echo "ncu rc=$?" >> $O/r02_s33.txt
echo done >> $O/r02_s33.txt
