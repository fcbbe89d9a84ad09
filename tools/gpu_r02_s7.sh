#!/bin/bash
# round 2, call 7: adaptive-bias PACK for long pairs, lean general path, warp-per-pair walker, window-edge tests:
# full GPU suite, fuzz (plain + wide), A/B on C2 / C3 / C5 against the build of call 6
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s7.log 2>&1; echo "pytest rc=$?" > $O/r02_s7.txt
timeout 200 python tools/fuzz_gpu.py 60 1234 > $O/fuzz_r02_s7.log 2>&1; echo "fuzz rc=$?" >> $O/r02_s7.txt
CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 60 4242 > $O/fuzz_r02_s7_wide.log 2>&1; echo "fuzz wide rc=$?" >> $O/r02_s7.txt
AB_WORKLOADS="C2 C3 C5" AB_STEPS=4 FUZZ_SECONDS=5 timeout 1200 tools/ab_variants.sh run pf12 r7 > $O/ab_r02_s7.txt 2>&1
echo "== C5 no_adapt" >> $O/r02_s7.txt
CLQ_NO_ADAPT=1 timeout 200 python bench.py --workload C5 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f" % (d["ms_per_step"], d["value"], d["gcups"]))' >> $O/r02_s7.txt

bash tools/canary_gpu.sh > $O/canary_r02_s7.log 2>&1; echo "canary rc=$?" >> $O/r02_s7.txt
echo done >> $O/r02_s7.txt
