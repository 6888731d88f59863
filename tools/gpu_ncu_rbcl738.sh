#!/bin/bash
# ncu --set full with source view of one rbcl738 evaluation's pruning launches (after warm-up)
mkdir -p gpurun_out
python tools/rbcl738_run.py 20 > gpurun_out/ncu738_plain.log 2>&1 || exit 1
# each evaluation = 6 prune launches; rbcl738_run does 1 + 10 warm-up evaluations first: skip 12 evaluations = 72 launches
ncu --set full --clock-control none --import-source on -k regex:prune_chains -s 72 -c 6 -f -o gpurun_out/rbcl738_prune \
    python tools/rbcl738_run.py 20 > gpurun_out/ncu738.log 2>&1
tail -3 gpurun_out/ncu738.log
ls -la gpurun_out/*.ncu-rep
