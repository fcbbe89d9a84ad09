#!/bin/bash
# round 2, call 30: the library with the AVX2 host packer (clq_pack2_host.cpp) -- packed-read GPU tests, smoke(), the full bench line
# (e2e_packed2 / e2e_api_packed2 with the faster packer), a short fuzz sweep.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s30.txt
grep -m1 "model name" /proc/cpuinfo >> $O/r02_s30.txt; grep -c ^processor /proc/cpuinfo >> $O/r02_s30.txt; grep -m1 -o "avx2" /proc/cpuinfo >> $O/r02_s30.txt
timeout -s KILL 300 python -m pytest tests/test_reads2bit.py tests/test_abi.py -q --timeout 200 > $O/pytest_r02_s30_reads2bit.log 2>&1; echo "reads2bit pytest (cpu + gpu) rc=$?" >> $O/r02_s30.txt; tail -2 $O/pytest_r02_s30_reads2bit.log >> $O/r02_s30.txt
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02_s30.log 2>&1; echo "smoke rc=$?" >> $O/r02_s30.txt; tail -1 $O/smoke_r02_s30.log >> $O/r02_s30.txt
(time timeout -s KILL 300 python bench.py > $O/bench_r02_s30.json 2> $O/bench_r02_s30.err) 2>> $O/r02_s30.txt; echo "bench rc=$?" >> $O/r02_s30.txt
timeout -s KILL 60 python tools/fuzz_gpu.py 40 20265 > $O/fuzz_r02_s30_seed20265.log 2>&1; tail -1 $O/fuzz_r02_s30_seed20265.log >> $O/r02_s30.txt
echo done >> $O/r02_s30.txt
