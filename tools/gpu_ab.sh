#!/bin/bash
# A/B of engine builds: tools/gpu_ab.sh <tag> <lib>...   (first lib "default" = the in-tree build)
mkdir -p gpurun_out
TAG=$1; shift
for lib in "$@"; do
  if [ "$lib" = default ]; then unset STROM_B200_LIB; name=default; else export STROM_B200_LIB=$lib; name=$(basename $(dirname $lib)); fi
  {
    echo "=== $name"
    python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us|fused call|logL -"
    PINVAR=0.2 python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us|fused call"
    python bench.py --taxa 1000 --patterns 250000 --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('250k ms/eval', d['ms_per_step'], 'prune ms', d['roofline']['prune_ms_per_eval'], 'clk', d['clocks'])"
  } >> gpurun_out/${TAG}_ab.log 2>&1
done
cat gpurun_out/${TAG}_ab.log
