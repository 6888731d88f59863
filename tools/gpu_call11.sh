#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
  echo "=== SB200_POLL=$v"
  SB200_POLL=$v python tools/facade_run.py | sed -n '2p;5p'
  SB200_POLL=$v python tools/rbcl738_run.py 1000 | tail -2
done
( timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
