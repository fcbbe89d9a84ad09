#!/bin/bash
# round 2, call 9: shared 96 GB traceback scratch, re-bias limit 30000; locate the wide-fuzz crash (verbose), C3 / C5 again,
# adaptive kernel at 2 CTAs/SM
cd "$(dirname "$0")/.."
O=gpurun_out
CLQ_FUZZ_VERBOSE=1 CLQ_FUZZ_WIDE=1 timeout 200 python tools/fuzz_gpu.py 60 4242 > $O/fuzz_r02_s9_wide.log 2>&1; echo "fuzz wide rc=$?" > $O/r02_s9.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02_s9.log 2>&1; echo "pytest rc=$?" >> $O/r02_s9.txt
AB_WORKLOADS="C3 C5" AB_STEPS=4 FUZZ_SECONDS=2 timeout 900 tools/ab_variants.sh run r9 a2 > $O/ab_r02_s9.txt 2>&1
echo "== C5 30k reads" >> $O/r02_s9.txt
timeout 200 python bench.py --workload C5 --reads 30000 --steps 3 --warmup 2 --no-cpu-baseline --no-live-peak --no-extra --no-api 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("ms %.3f reads/s %.4g gcups %.1f sub_batches %s retries %s" % (d["ms_per_step"], d["value"], d["gcups"], d["config"].get("sub_batches"), d["config"].get("pack_retries")))' >> $O/r02_s9.txt
echo done >> $O/r02_s9.txt
