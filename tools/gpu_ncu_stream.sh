#!/bin/bash
# ncu --set full (source view) of the five pruning launches of one 1000 x 250k evaluation
mkdir -p gpurun_out
python bench.py --taxa 1000 --patterns 250000 --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/plain_stream.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:prune_chains -s 10 -c 5 -f -o gpurun_out/stream_prune \
    python bench.py --taxa 1000 --patterns 250000 --steps 2 --warmup 3 --no-cpu --no-f32 > gpurun_out/ncu_stream.log 2>&1
tail -2 gpurun_out/ncu_stream.log | cut -c1-300; ls -la gpurun_out/stream_prune.ncu-rep
