#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q ) > gpurun_out/c4_fullsize.log 2>&1
for v in 1 2 1 2; do
  {
    echo "=== SB200_SMALL_VARIANT=$v"
    SB200_SMALL_VARIANT=$v python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us|fused call|logL -"
    SB200_SMALL_VARIANT=$v PINVAR=0.2 python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us|fused call"
  } >> gpurun_out/c4_ab.log 2>&1
done
( SB200_SMALL_VARIANT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > gpurun_out/c4_parity_v2.log 2>&1
tail -5 gpurun_out/c4_fullsize.log; cat gpurun_out/c4_ab.log; tail -3 gpurun_out/c4_parity_v2.log
