#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g4_pytest_all.log 2>&1; echo "pytest rc $?" >> gpurun_out/g4_pytest_all.log; tail -n 4 gpurun_out/g4_pytest_all.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err; tail -n 2 gpurun_out/g4_bench.err
python - <<PY
import json
for l in open("gpurun_out/g4_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value %.4e" % d["value"], "ms", round(d["ms_per_step"], 3), "sweep", d["sweep"]["wall_s"], d["sweep"]["solutions_in_flight_per_gpu"], "sharded", d["sharded"]["wall_s"], "solution", d["solution"]["em_wall_s"], d["solution"]["viterbi_wall_s"])
PY
