#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "fused_r_glue or r_call_glue" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest.log
tail -15 gpurun_out/r02c_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -5 gpurun_out/r02c_bench.err
cat gpurun_out/r02c_bench.json | cut -c1-300
