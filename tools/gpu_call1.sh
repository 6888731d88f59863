#!/bin/bash
# round-1 re-entry call 1: parity on all category counts, rbcl738 C=4 / C=5 timing, latency microbenchmarks, step traces
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
python tools/rbcl738_run.py 500 > gpurun_out/c1_rbcl738_c4.log 2>&1
PINVAR=0.2 python tools/rbcl738_run.py 500 > gpurun_out/c1_rbcl738_c5.log 2>&1
./tools/latbench > gpurun_out/c1_latbench.log 2>&1
for lv in 0 1 3 5; do
  STROM_B200_LIB=tools/_trace/libstrom_b200.so SB200_TRACE_LEVEL=$lv python tools/trace_rbcl738.py > gpurun_out/c1_trace_l$lv.log 2>&1
done
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_rbcl738_c4.log gpurun_out/c1_rbcl738_c5.log gpurun_out/c1_latbench.log
