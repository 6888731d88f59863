#!/bin/bash
mkdir -p gpurun_out
r() { python tools/rbcl738_run.py 500 2>&1 | grep "device us" ; }
{
echo "== materialised default"; r; PINVAR=0.2 r
export TRANSIENT=1
echo "== transient default"; r; PINVAR=0.2 r
for w in 600 1200 4800 9600; do echo "== wide min warps $w"; SB200_WIDE_MIN_WARPS=$w r; done
for w in 150 300 1200; do echo "== narrow min warps $w"; SB200_NARROW_MIN_WARPS=$w r; done
for s in 1184 3552 7104; do echo "== slots $s"; SB200_SLOTS=$s r; done
for s in 592 2368; do echo "== tips slots $s"; SB200_TIPS_SLOTS=$s r; done
echo "== no pdl"; SB200_PDL=0 r
echo "== batch 8/16 transient"; python tools/rbcl738_batch.py 8 200 2>&1 | grep "us/batch"; python tools/rbcl738_batch.py 16 200 2>&1 | grep "us/batch"
for w in 1200 9600 40000; do echo "== batch wide min warps $w"; SB200_WIDE_MIN_WARPS=$w python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; done
for s in 1184 2368 7104; do echo "== batch slots $s"; SB200_SLOTS=$s python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; done
} > gpurun_out/sweep8.log 2>&1
cat gpurun_out/sweep8.log
