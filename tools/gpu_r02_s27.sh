#!/bin/bash
# round 2, call 27 (2 GPUs): span dispatcher A/B on a sorted C5 stream (tools/span_experiment.py)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 420 python tools/span_experiment.py 30000 > $O/span_experiment_r02_s27.txt 2>&1; echo "rc=$?" >> $O/span_experiment_r02_s27.txt
