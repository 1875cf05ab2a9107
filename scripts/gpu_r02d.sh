#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d_bench2.json 2> gpurun_out/r02d_bench2.err; tail -5 gpurun_out/r02d_bench2.err
cat gpurun_out/r02d_bench2.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-budget-s 30 > gpurun_out/r02d_ref2.json 2> gpurun_out/r02d_ref2.err; tail -3 gpurun_out/r02d_ref2.err
