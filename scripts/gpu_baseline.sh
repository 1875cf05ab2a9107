#!/bin/bash
# r02 baseline: gpu tests, bench, launch list, full ncu capture of the E-step kernels (run under gpurun, 1 GPU)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 2 > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^fb2_" -c 14 -o gpurun_out/r02_fb2_full -f \
  python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 1 > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_pytest.log; cat gpurun_out/r02_bench.json; cat gpurun_out/r02_prof_plain.log
