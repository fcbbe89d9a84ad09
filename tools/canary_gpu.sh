#!/bin/bash
# Window canary (VERDICT r1 weak #1a): run the fuzz sweeps and the bench samples on a -DCLQ_PACK_CANARY=1 build, which tracks every
# value the s16x2 kernels store and counts the tasks that left [64, 32767] without being covered by the retry pass.  Must print
# "window violations 0".   here:  tools/ab_variants.sh build canary="-DCLQ_LEAN=1 -DCLQ_PACK_CANARY=1"
#                          box:   gpurun --timeout 900 -- 'bash tools/canary_gpu.sh > gpurun_out/canary.log 2>&1'
cd "$(dirname "$0")/.."
cp clique_b200/libclq.so tools/_v/.in_tree.so
trap 'cp tools/_v/.in_tree.so clique_b200/libclq.so' EXIT
cp tools/_v/libclq_canary.so clique_b200/libclq.so
CLQ_FUZZ_NO_PACK_P=0.05 timeout 150 python tools/fuzz_gpu.py ${CANARY_SECONDS:-90} 31337
CLQ_FUZZ_WIDE=1 CLQ_FUZZ_NO_PACK_P=0.05 CLQ_FUZZ_MODES=fixed,fixed,fixed,exhaustive timeout 250 python tools/fuzz_gpu.py ${CANARY_SECONDS:-90} 271828
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "window_edge or adaptive or c5_sample or c3_sample or c2_sample" 2>&1 | tail -2
timeout 200 python - <<'PY'
import ctypes, sys
sys.path.insert(0, ".")
import bench
from clique_b200 import load_library
import argparse
for wl, n in (("C2", 200000), ("C3", 20000), ("C5", 8000)):
    args = argparse.Namespace()
    r = bench.run_config(args, wl, n, "", False, False, 1, 1, 2, 0, 0, 1, lambda: None)
    print(wl, "reads", n, "ok", r["n_ok"], "variant", r["variant"])
lib = load_library()
chk, vio = ctypes.c_ulonglong(), ctypes.c_ulonglong()
print("rc", lib.clq_debug_canary(ctypes.byref(chk), ctypes.byref(vio)), "canary: s16x2 tasks checked", chk.value, "window violations", vio.value)
PY
