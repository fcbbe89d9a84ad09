#!/bin/bash
# round 2, call 6: walker prefetch distance A/B (C2, C5), packed k-mer vote kernel (quick-search tests + C4 quick), e2e_api with ramp-up
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s6.log 2>&1; echo "pytest rc=$?" > $O/r02_s6.txt
AB_WORKLOADS="C2 C5" AB_STEPS=4 FUZZ_SECONDS=10 timeout 900 tools/ab_variants.sh run pf12 pf0 pf24 > $O/ab_r02_s6_walkpf.txt 2>&1
echo "== C4 quick" >> $O/r02_s6.txt
timeout 120 python bench.py --workload C4 --search quick --reads 400000 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f kernel_ms %.3f" % (d["ms_per_step"], d["value"], d["gcups"], d["roofline"]["kernel_ms"]))' >> $O/r02_s6.txt
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 5 > $O/bench_r02_s6_api.json 2> $O/bench_r02_s6_api.err; echo "bench rc=$?" >> $O/r02_s6.txt
echo done >> $O/r02_s6.txt
