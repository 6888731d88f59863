#!/bin/bash
mkdir -p gpurun_out
b() { python bench.py --taxa 1000 --patterns $1 --steps $2 --warmup 3 --no-cpu --no-f32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 ms', d['ms_per_step'], 'p50', d['ms_per_step_p50'], 'sched_frac', d['roofline'].get('scheduled_frac'), d['clocks']['sm_mhz'])"; }
r738() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
{
echo "== default"; b 250000 5; b 1000000 10; r738
for v in $(ls tools/_var); do echo "== lib $v"; export STROM_B200_LIB=$PWD/tools/_var/$v/libstrom_b200.so; b 250000 5; b 1000000 10; r738; python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; unset STROM_B200_LIB; done
for s in 592 1184; do echo "== slots $s"; SB200_SLOTS=$s b 250000 5; done
} > gpurun_out/sweep5.log 2>&1
cat gpurun_out/sweep5.log
