#!/bin/bash
# Quick sweep of one engine switch on single rbcl738 evaluations (GTR+G4, transient, GTR+I+G4): usage gpu_sweep2.sh VAR v1 v2 ...
mkdir -p gpurun_out
var=$1; shift
r() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
for v in "$@"; do
  echo "== $var=$v"; export $var=$v
  r; TRANSIENT=1 r; PINVAR=0.2 r
done 2>&1 | tee -a gpurun_out/sweep2b.log
