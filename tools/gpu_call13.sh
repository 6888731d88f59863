#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_host.py -x -q ) 2>&1 | tail -3
python bench.py --taxa 1000 --patterns 100000 --steps 3 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo rc $?
python -c "
import json; d=json.load(open('gpurun_out/c13_bench.json'))
for b in ('rbcl738','rbcl738_i_g4'):
    print(b, {k: round(v,1) for k,v in d[b].items() if k.startswith('evals')})
"
