#!/bin/bash
# Round-2 baseline evidence of the HEAD kernels, one gpurun call (1 GPU):
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02_baseline.sh'
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_box.txt 2>&1
nproc >> $O/r02_box.txt
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s1.log 2>&1; echo "pytest rc=$?" >> $O/r02_box.txt
timeout 300 python bench.py > $O/bench_r02_s1_C2.json 2> $O/bench_r02_s1_C2.err; echo "bench rc=$?" >> $O/r02_box.txt
timeout 60 tools/int_peak > $O/int_peak_r02.jsonl 2>&1
timeout 600 tools/ab_variants.sh run base ep > $O/ab_r02_s1_ep.txt 2>&1
CMD="python bench.py --reads 400000 --steps 2 --warmup 1 --no-cpu-baseline --no-live-peak"
timeout 200 $CMD > $O/plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_s1.csv $CMD > $O/ncu_l.log 2>&1
CMD2="python bench.py --reads 400000 --steps 1 --warmup 1 --no-cpu-baseline --no-live-peak"
timeout 200 $CMD2 > $O/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pack_kernel|walk_kernel' -c 4 -f -o $O/prof_r02_s1 $CMD2 > $O/ncu_f.log 2>&1
echo done >> $O/r02_box.txt
