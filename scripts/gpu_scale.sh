#!/bin/bash
# bench at N = 8 and N = 4 (run under gpurun --gpus 8)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; tail -3 gpurun_out/r02_bench_n$n.err
done
python - <<'PY'
import json
for n in (8, 4):
    for l in open(f"gpurun_out/r02_bench_n{n}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(n, "value", d["value"], "ms", d["ms_per_step"], "sweep", d.get("sweep", {}).get("wall_s"), d.get("sweep", {}).get("lpt_load_K"),
                  "sharded", d.get("sharded", {}).get("wall_s"), d.get("sharded", {}).get("estep_device_ms_max_rank"))
PY
