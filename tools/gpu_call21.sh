#!/bin/bash
for v in 0 1 0 1; do
  echo "=== SB200_DIRECT_PARAMS=$v"
  SB200_DIRECT_PARAMS=$v python tools/facade_run.py | sed -n '2p;5p'
  SB200_DIRECT_PARAMS=$v python tools/rbcl738_run.py 1000 | tail -3
done
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q ) 2>&1 | tail -2
