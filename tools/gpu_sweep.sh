#!/bin/bash
mkdir -p gpurun_out
r() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
bt() { python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; }
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
{
echo "== default"; r; TRANSIENT=1 r; PINVAR=0.2 r; bt; TRANSIENT=1 bt; PINVAR=0.2 bt
for w in 2368 4736 9472 20000 100000; do echo "== wide min warps $w"; SB200_WIDE_MIN_WARPS=$w r; SB200_WIDE_MIN_WARPS=$w TRANSIENT=1 r; SB200_WIDE_MIN_WARPS=$w bt; done
for w in 1184 4736; do echo "== wide2 min warps $w"; SB200_WIDE2_MIN_WARPS=$w r; SB200_WIDE2_MIN_WARPS=$w bt; done
python bench.py --taxa 1000 --patterns 1000000 --steps 8 --warmup 3 --no-cpu --no-f32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('1M ms', d['ms_per_step'], 'sched_frac', d['roofline'].get('scheduled_frac'), d['clocks']['sm_mhz'])"
} > gpurun_out/sweep10.log 2>&1
cat gpurun_out/sweep10.log
