#!/bin/bash
# A/B of a tuning switch: tools/gpu_ab_env.sh VAR v1 v2 ...   (rbcl738 device / fused-call / facade timing per value)
VAR=$1; shift
for v in "$@"; do
  echo "=== $VAR=$v"
  env $VAR=$v python tools/rbcl738_run.py 1000 | grep -E "device us|enqueue"
  env $VAR=$v PINVAR=0.2 python tools/rbcl738_run.py 1000 | grep -E "device us"
  env $VAR=$v python tools/facade_run.py | sed -n '2p;5p'
done
