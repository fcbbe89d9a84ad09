#!/bin/bash
# round 2, call 31 (last): ncu --set full of the packed-read expansion kernels (north_star: "achieved HBM GB/s for the ... packed-read
# traffic"), then the final evidence on HEAD: full GPU suite and the complete bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s31.txt
timeout -s KILL 120 python tools/unpack_probe.py > $O/unpack_probe_r02.txt 2>&1; echo "probe rc=$?" >> $O/r02_s31.txt; cat $O/unpack_probe_r02.txt >> $O/r02_s31.txt
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:'unpack2_kernel|patch2_kernel' -c 4 -f -o $O/prof_r02_unpack2 python tools/unpack_probe.py > $O/ncu_unpack2.log 2>&1; echo "ncu rc=$?" >> $O/r02_s31.txt
CMD="python bench.py --reads 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout -s KILL 120 $CMD > $O/plain_final2.log 2>&1 && \
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_final2_C2.csv $CMD > $O/ncu_l_final2.log 2>&1
echo "launch list rc=$?" >> $O/r02_s31.txt
(time timeout -s KILL 300 python bench.py > $O/bench_r02_final2.json 2> $O/bench_r02_final2.err) 2>> $O/r02_s31.txt; echo "bench rc=$?" >> $O/r02_s31.txt
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 > $O/pytest_gpu_r02_final2.log 2>&1; echo "pytest rc=$?" >> $O/r02_s31.txt; tail -3 $O/pytest_gpu_r02_final2.log >> $O/r02_s31.txt
echo done >> $O/r02_s31.txt
