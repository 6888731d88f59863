#!/bin/bash
# One GPU call: parity tests, rbcl738 timing (GTR+G4 and GTR+I+G4), a 250k-pattern bench line.  Output under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/rbcl738_run.py 500 > gpurun_out/rbcl738_c4.log 2>&1; tail -4 gpurun_out/rbcl738_c4.log
PINVAR=0.2 python tools/rbcl738_run.py 500 > gpurun_out/rbcl738_c5.log 2>&1; tail -4 gpurun_out/rbcl738_c5.log
python bench.py --taxa 1000 --patterns 250000 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_250k.log 2>&1; tail -2 gpurun_out/bench_250k.log | cut -c1-1500
