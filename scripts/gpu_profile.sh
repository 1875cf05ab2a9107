#!/bin/bash
# round-2 profile capture (1 GPU): launch list of a bench run + ncu --set full of one whole E-step and of one Viterbi decode
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --no-sharded > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sweep --no-sharded > gpurun_out/r02_ncu_list.log 2>&1
python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 3 > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^(fb2_|reduce_)" -c 40 -o gpurun_out/r02_estep_full -f \
  python scripts/prof_estep.py --snps 1000000 --chunk 1024 --warm 192 --steps 3 > gpurun_out/r02_ncu_full.log 2>&1
python scripts/vit_bench.py --clusters 3 --modes 4 > gpurun_out/r02_vit_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"viterbi_redux" -c 1 -o gpurun_out/r02_viterbi_full -f \
  python scripts/vit_bench.py --clusters 3 --modes 4 > gpurun_out/r02_ncu_vit.log 2>&1
# digest on the box: the .ncu-rep files are too large to bring back
mkdir -p gpurun_out/prof
python scripts/make_profile_summary.py r02 gpurun_out/r02_estep_full.ncu-rep gpurun_out/r02_launches.csv gpurun_out/prof
ncu -i gpurun_out/r02_viterbi_full.ncu-rep --page raw --csv > gpurun_out/prof/r02_ncu_viterbi_redux_raw.csv
python scripts/ncu_digest.py gpurun_out/r02_viterbi_full.ncu-rep gpurun_out/prof/r02_ncu_viterbi_summary.json
cp gpurun_out/r02_launches.csv gpurun_out/prof/
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/r02_prof_plain.log gpurun_out/r02_vit_plain.log
