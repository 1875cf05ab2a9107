#!/bin/bash
# E-step time against chunk length / warm-up length (bench workload)
for c in 256 384 512 768 1024; do for w in 96 192; do
  echo "chunk $c warm $w"; python bench.py --steps 10 --warmup 3 --no-em --no-cpu-baseline --chunk $c --warm $w | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(' ms_per_step',round(d['ms_per_step'],3),'rounds',d['config']['fwd_rounds'],'launches',d['gpu_launches'],'r0',round(d['roofline']['fwd_kernel']['launch_ms'],3),'final',round(d['roofline']['launch_ms'],3))"
done; done 2>&1 | tee gpurun_out/r02_chunk_sweep.log
