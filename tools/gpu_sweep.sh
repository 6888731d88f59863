#!/bin/bash
# A/B of one engine switch on rbcl738 (single evaluation, 8 batched chains; GTR+G4 and GTR+I+G4): usage gpu_sweep.sh VAR v1 v2 ...
mkdir -p gpurun_out
var=$1; shift
r() { python tools/rbcl738_run.py 500 2>&1 | grep "device us\|enqueue" ; }
bt() { python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us\|fused"; }
for v in "$@"; do
  echo "== $var=$v"; export $var=$v
  r; TRANSIENT=1 r; PINVAR=0.2 r; bt; TRANSIENT=1 bt; PINVAR=0.2 bt
done 2>&1 | tee gpurun_out/sweep.log
