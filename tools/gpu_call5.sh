#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/c5_ab.log
for v in 0 1 0 1; do
  {
    echo "=== SB200_TIPS_KERNEL=$v"
    SB200_TIPS_KERNEL=$v python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us|logL -"
    SB200_TIPS_KERNEL=$v PINVAR=0.2 python tools/rbcl738_run.py 1000 2>&1 | grep -E "device us"
    SB200_TIPS_KERNEL=$v python bench.py --taxa 1000 --patterns 250000 --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('250k ms/eval', d['ms_per_step'], 'prune ms', d['roofline']['prune_ms_per_eval'], 'logL', d['logL'], 'clk', d['clocks'])"
  } >> gpurun_out/c5_ab.log 2>&1
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c5_pytest.log 2>&1
cat gpurun_out/c5_ab.log; tail -5 gpurun_out/c5_pytest.log
