#!/bin/bash
# round 2, call 32: ncu --set full of unpack2_kernel / patch2_kernel (the probe of call 31 asked for slot stats without a launch and exited 1)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s32.txt
timeout -s KILL 100 python tools/unpack_probe.py > $O/unpack_probe_r02.txt 2>&1; echo "probe rc=$?" >> $O/r02_s32.txt; cat $O/unpack_probe_r02.txt >> $O/r02_s32.txt
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:'unpack2_kernel|patch2_kernel' -c 4 -f -o $O/prof_r02_unpack2 python tools/unpack_probe.py > $O/ncu_unpack2.log 2>&1; echo "ncu rc=$?" >> $O/r02_s32.txt
echo done >> $O/r02_s32.txt
