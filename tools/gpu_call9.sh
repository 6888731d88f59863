#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c9_pytest.log 2>&1
python tools/rbcl738_parts.py > gpurun_out/c9_parts.log 2>&1
python tools/rbcl738_run.py 1000 > gpurun_out/c9_run.log 2>&1
PINVAR=0.2 python tools/rbcl738_run.py 1000 >> gpurun_out/c9_run.log 2>&1
tail -5 gpurun_out/c9_pytest.log; cat gpurun_out/c9_parts.log gpurun_out/c9_run.log
