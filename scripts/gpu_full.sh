#!/bin/bash
# the driver's round-end sequence on one GPU: gpu tests, smoke, both bench arms
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/full_pytest.log; tail -4 gpurun_out/full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/full_ref.json 2> gpurun_out/full_ref.err; cut -c1-200 gpurun_out/full_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -3 gpurun_out/full_bench.err; cut -c1-300 gpurun_out/full_bench.json
