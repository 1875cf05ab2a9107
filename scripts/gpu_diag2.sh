#!/bin/bash
mkdir -p gpurun_out
echo "=== derived remainder K=75" > gpurun_out/diag6.log
DIAG_CFG="1024,192,1e-8" TITAN_TRACE=2 python scripts/diag_rounds.py 2>&1 | grep "^iter\|chunk runs" | awk 'NR>6' | uniq >> gpurun_out/diag6.log 2>&1
echo "=== derived remainder K=125" >> gpurun_out/diag6.log
DIAG_N=2000000 DIAG_Z=5 DIAG_CFG="1024,1024,1e-8" TITAN_TRACE=2 python scripts/diag_rounds.py 2>&1 | grep "^iter\|chunk runs" | awk 'NR>7' | uniq >> gpurun_out/diag6.log 2>&1
cat gpurun_out/diag6.log | cut -c1-250
python -m pytest tests -m gpu -x -q > gpurun_out/g2_pytest_all.log 2>&1; echo "pytest rc $?" >> gpurun_out/g2_pytest_all.log; tail -n 15 gpurun_out/g2_pytest_all.log
