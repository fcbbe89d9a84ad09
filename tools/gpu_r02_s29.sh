#!/bin/bash
# round 2, call 29: the 2-bit packed upload (clq_pack2 / clq_submit_packed2, unpack2_kernel) -- its GPU tests, the full GPU suite on the
# rebuilt library, smoke(), the bench line with e2e_packed2 / e2e_api_packed2, a fuzz sweep in which every third batch ships packed.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s29.txt
timeout -s KILL 300 python -m pytest tests/test_reads2bit.py -m gpu -q --timeout 200 > $O/pytest_gpu_r02_reads2bit.log 2>&1; echo "reads2bit pytest rc=$?" >> $O/r02_s29.txt; tail -3 $O/pytest_gpu_r02_reads2bit.log >> $O/r02_s29.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02_s29.log 2>&1; echo "smoke rc=$?" >> $O/r02_s29.txt; tail -1 $O/smoke_r02_s29.log >> $O/r02_s29.txt
(time timeout -s KILL 300 python bench.py > $O/bench_r02_s29.json 2> $O/bench_r02_s29.err) 2>> $O/r02_s29.txt; echo "bench rc=$?" >> $O/r02_s29.txt
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 --deselect tests/test_reads2bit.py > $O/pytest_gpu_r02_s29.log 2>&1; echo "pytest rc=$?" >> $O/r02_s29.txt; tail -3 $O/pytest_gpu_r02_s29.log >> $O/r02_s29.txt
timeout -s KILL 70 python tools/fuzz_gpu.py 50 20264 > $O/fuzz_r02_s29_seed20264.log 2>&1; tail -1 $O/fuzz_r02_s29_seed20264.log >> $O/r02_s29.txt
echo done >> $O/r02_s29.txt
