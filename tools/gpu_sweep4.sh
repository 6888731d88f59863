#!/bin/bash
mkdir -p gpurun_out
r738() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
b250() { python bench.py --taxa 1000 --patterns 250000 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('250k ms', d['ms_per_step'], 'sched_frac', d['roofline'].get('scheduled_frac'))"; }
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
echo "== default"; r738; PINVAR=0.2 r738; b250
for v in $(ls tools/_var); do echo "== lib $v"; export STROM_B200_LIB=$PWD/tools/_var/$v/libstrom_b200.so; b250; unset STROM_B200_LIB; done
echo "== all quad"; SB200_VARIANT=3 r738
echo "== all narrow"; SB200_VARIANT=2 r738
echo "== all wide"; SB200_VARIANT=1 r738
for w in 300 1200 2400; do echo "== narrow min warps $w"; SB200_NARROW_MIN_WARPS=$w r738; done
for w in 1200 4800 100000; do echo "== wide min warps $w"; SB200_WIDE_MIN_WARPS=$w r738; done
for k in 8; do python tools/rbcl738_batch.py $k 200 2>&1 | tail -3; done
echo "== PINVAR"; PINVAR=0.2 python tools/rbcl738_batch.py 8 200 2>&1 | tail -2
} > gpurun_out/sweep4.log 2>&1
cat gpurun_out/sweep4.log
python tools/rbcl738_run.py 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/r2_rbcl738_levels.csv python tools/rbcl738_run.py 3 > gpurun_out/ncu1.log 2>&1
python tools/rbcl738_batch.py 8 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 150 -c 12 --csv --log-file gpurun_out/r2_rbcl738_batch8_levels.csv python tools/rbcl738_batch.py 8 3 > gpurun_out/ncu2.log 2>&1
echo done
