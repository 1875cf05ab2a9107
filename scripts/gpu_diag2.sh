#!/bin/bash
mkdir -p gpurun_out
for w in 16 32; do
echo "=== window $w K=125"
TITAN_PROBE_WINDOW=$w DIAG_N=2000000 DIAG_Z=5 DIAG_CFG="1024,1024,1e-8" TITAN_TRACE=2 python scripts/diag_rounds.py 2>&1 | grep "^iter\|chunk runs" | awk 'NR>7' | uniq
done > gpurun_out/diag5.log 2>&1
echo "=== window 32 K=75" >> gpurun_out/diag5.log
TITAN_PROBE_WINDOW=32 DIAG_CFG="1024,192,1e-8" TITAN_TRACE=2 python scripts/diag_rounds.py 2>&1 | grep "^iter\|chunk runs" | awk 'NR>6' | uniq >> gpurun_out/diag5.log 2>&1
cat gpurun_out/diag5.log | cut -c1-250
