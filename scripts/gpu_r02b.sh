#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -5 gpurun_out/r02b_bench.err
cat gpurun_out/r02b_bench.json
python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 3 > gpurun_out/r02b_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^(fb2_|reduce_)" -c 30 -o gpurun_out/r02b_estep_full -f \
  python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 3 > gpurun_out/r02b_ncu_full.log 2>&1
tail -3 gpurun_out/r02b_ncu_full.log
