#!/bin/bash
# round 2, call 34: sanity run of the clean-rebuilt final binaries (smoke + the golden / C2-sample / packed-read GPU tests)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r02_s34.txt
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02_s34.log 2>&1; echo "smoke rc=$?" >> $O/r02_s34.txt; tail -1 $O/smoke_r02_s34.log >> $O/r02_s34.txt
timeout -s KILL 80 python -m pytest tests/test_reads2bit.py tests/test_gpu_parity.py -m gpu -q --timeout 60 -k "reads2bit or goldens or c2_sample or packed or unpack" > $O/pytest_gpu_r02_s34.log 2>&1; echo "pytest rc=$?" >> $O/r02_s34.txt; tail -2 $O/pytest_gpu_r02_s34.log >> $O/r02_s34.txt
echo done >> $O/r02_s34.txt
