#!/bin/bash
mkdir -p gpurun_out
b() { python bench.py --taxa 1000 --patterns $1 --steps $2 --warmup 3 --no-cpu --no-f32 $3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 $3 ms', d['ms_per_step'], 'p50', d['ms_per_step_p50'], 'sched_frac', d['roofline'].get('scheduled_frac'), d['clocks']['sm_mhz'])"; }
one() { b 250000 5; b 250000 5 --transient; b 1000000 8; b 1000000 8 --transient; python tools/rbcl738_run.py 500 2>&1 | grep "device us"; TRANSIENT=1 python tools/rbcl738_run.py 500 2>&1 | grep "device us"; python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; TRANSIENT=1 python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; }
{
echo "== default"; one
for v in $(ls tools/_var); do echo "== lib $v"; export STROM_B200_LIB=$PWD/tools/_var/$v/libstrom_b200.so; python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1; one; unset STROM_B200_LIB; done
} > gpurun_out/sweep9.log 2>&1
cat gpurun_out/sweep9.log
