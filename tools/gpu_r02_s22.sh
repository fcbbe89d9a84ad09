#!/bin/bash
# round 2, call 22 (2 GPUs): the bench line under torchrun at N=2 as the driver launches it (weak-scaling replicas + the `sharded`
# block: rank 0 drives both GPUs through ShardedAligner::align_reads_span over one C5 stream), the reference arm at N=2, and the
# two multi-GPU tests that are skipped on 1-GPU boxes.
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r02_s22.txt
nvidia-smi -L >> $O/r02_s22.txt
(time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_r02_s22_n2.json 2> $O/bench_r02_s22_n2.err) 2>> $O/r02_s22.txt; echo "bench n2 rc=$?" >> $O/r02_s22.txt
(time timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_r02_s22_ref_n2.json 2> $O/bench_r02_s22_ref_n2.err) 2>> $O/r02_s22.txt; echo "reference n2 rc=$?" >> $O/r02_s22.txt
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 -k "two_gpus or all_gpus or shard" > $O/pytest_gpu_r02_s22.log 2>&1; echo "pytest rc=$?" >> $O/r02_s22.txt; tail -3 $O/pytest_gpu_r02_s22.log >> $O/r02_s22.txt
echo done >> $O/r02_s22.txt
