#!/bin/bash
# A/B of engine builds (tools/_var/<name>/libstrom_b200.so, see build_variant.sh) on rbcl738: usage gpu_ab.sh name...
mkdir -p gpurun_out
r() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
bt() { python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; }
one() { r; TRANSIENT=1 r; PINVAR=0.2 r; bt; TRANSIENT=1 bt; }
{ echo "== default"; one
for v in "$@"; do echo "== lib $v"; export STROM_B200_LIB=$PWD/tools/_var/$v/libstrom_b200.so; one; unset STROM_B200_LIB; done; } 2>&1 | tee gpurun_out/ab.log
