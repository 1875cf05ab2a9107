#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gaussian.py -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g1_pytest.log; tail -n 40 gpurun_out/g1_pytest.log
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_gaussian.py > gpurun_out/g1_pytest_all.log 2>&1; echo "pytest rc $?" >> gpurun_out/g1_pytest_all.log; tail -n 6 gpurun_out/g1_pytest_all.log
