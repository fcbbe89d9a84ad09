#!/bin/bash
# round 2, call 4: e2e_api after the claim-policy fix, launch lists of C2 / C4 quick / C5, ncu --set full of the restructured C2 kernels
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 5 > $O/bench_r02_s4_api.json 2> $O/bench_r02_s4_api.err; echo "bench rc=$?" > $O/r02_s4.txt
for W in "C2 --reads 400000" "C4 --search quick --reads 200000" "C5 --reads 12000"; do
  set -- $W; name=$1
  CMD="python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
  timeout 200 $CMD > $O/plain_$name.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r02_s4_$name.csv $CMD > $O/ncu_l_$name.log 2>&1
done
CMD2="python bench.py --reads 400000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak --no-extra --no-api"
timeout 200 $CMD2 > $O/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pack_kernel|walk_kernel' -c 4 -f -o $O/prof_r02_s4 $CMD2 > $O/ncu_f.log 2>&1
echo done >> $O/r02_s4.txt
