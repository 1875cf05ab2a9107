#!/bin/bash
# bench at N = 8 and N = 4 (run under gpurun --gpus 8), or N = 2 + the sharded test (gpurun --gpus 2)
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 8 ]; then LIST="8 4"; else LIST="2"; python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2; fi
for n in $LIST; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; tail -3 gpurun_out/r02_bench_n$n.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n$n.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print($n, "value %.4e" % d["value"], "ms", round(d["ms_per_step"], 3), "rank_ms", [round(x, 3) for x in d["engine"]["rank_ms_per_step"]])
        print("  sweep", d.get("sweep", {}).get("wall_s"), d.get("sweep", {}).get("lpt_load_K"))
        print("  sharded", {k: d["sharded"][k] for k in ("wall_s", "estep_device_ms_max_rank", "estep_compute_ms_max_rank")})
PY
done
