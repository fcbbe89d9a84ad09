#!/bin/bash
# A/B of kernel build variants in ONE gpurun call (same box, same clocks), the way the round-1 experiments were measured
# (profiles/ab_r01_s3_ext2_relu.txt):
#   tools/ab_variants.sh build base= ep=-DCLQ_PACK_EXT2_VIA_EP=1     # here: one libclq_<name>.so per variant under tools/_v/
#   gpurun --timeout 240 -- 'tools/ab_variants.sh run base ep'       # on the GPU box: focused fuzz + C2 bench per variant
# tools/_v/ is scratch (git-ignored); the in-tree clique_b200/libclq.so is restored after `run`.
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
case "$mode" in
build)
    mkdir -p tools/_v
    for v in "$@"; do
        name=${v%%=*}; flags=${v#*=}
        echo "== $name: $flags"
        (cd clique_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
            -shared -o ../../tools/_v/libclq_$name.so clq_api.cu clq_pack2_host.cpp -lcudart)
    done ;;
run)
    cp clique_b200/libclq.so tools/_v/.in_tree.so
    trap 'cp tools/_v/.in_tree.so clique_b200/libclq.so' EXIT
    for name in "$@" "$1"; do   # the first variant twice: run-to-run noise
        cp tools/_v/libclq_$name.so clique_b200/libclq.so
        echo "== $name"
        CLQ_FUZZ_MODES=fixed,fixed,exhaustive,quick CLQ_FUZZ_NO_PACK_P=0.1 timeout 60 python tools/fuzz_gpu.py ${FUZZ_SECONDS:-25} 4711 2>&1 | tail -2
        for wl in ${AB_WORKLOADS:-C2}; do
        timeout 120 python bench.py --workload $wl --steps ${AB_STEPS:-6} --warmup 3 --no-cpu-baseline --no-live-peak --no-extra 2>/dev/null | python -c '
import sys, json
d = json.loads(sys.stdin.readline())
print("%s ms_per_step %.3f  reads/s %.4g  gcups %.1f  e2e %.4g  dp_kernel_ms %.3f  ok_reads %d  sub_batches %s  pack_retries %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["gcups"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["config"]["status_ok_reads"], d["config"].get("sub_batches"), d["config"].get("pack_retries")))' $wl
        done
    done ;;
*) echo "usage: $0 build name=flags... | run name..."; exit 2 ;;
esac
