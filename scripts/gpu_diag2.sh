#!/bin/bash
mkdir -p gpurun_out
python scripts/sweep_breakdown.py 2>&1 | tail -n 17 | tee gpurun_out/sweep_breakdown.log
python -m pytest tests -m gpu -x -q > gpurun_out/g3_pytest_all.log 2>&1; echo "pytest rc $?" >> gpurun_out/g3_pytest_all.log; tail -n 8 gpurun_out/g3_pytest_all.log
