#!/bin/bash
mkdir -p gpurun_out
b() { python bench.py --taxa 1000 --patterns $1 --steps $2 --warmup 3 --no-cpu --no-f32 $3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 $3 ms', d['ms_per_step'], 'p50', d['ms_per_step_p50'], 'sched_frac', d['roofline'].get('scheduled_frac'), 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'], 'logL', d['logL'])"; }
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transient or shape or batched" 2>&1 | tail -3
{
b 250000 5; b 250000 5 --transient; b 1000000 10; b 1000000 10 --transient
python tools/rbcl738_run.py 500 2>&1 | grep "device us\|logL"; TRANSIENT=1 python tools/rbcl738_run.py 500 2>&1 | grep "device us\|logL"
python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; TRANSIENT=1 python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us\|rel diff"
PINVAR=0.2 TRANSIENT=1 python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us\|rel diff"
} > gpurun_out/sweep6.log 2>&1
cat gpurun_out/sweep6.log
