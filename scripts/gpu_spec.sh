python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -x -q -k "estep or em_ or config4 or chunked or full_size or edge or bitwise or sharded or comm or kernel_variants or readme or sweep" 2>&1 | tail -6
for sp in 0 1; do echo "SPEC_FINAL $sp"; TITAN_SPEC_FINAL=$sp python bench.py --steps 10 --warmup 3 --no-em --no-sweep --no-sharded --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['engine']['cycles_per_step'], {k: round(v,3) for k,v in d['roofline']['phases_ms'].items()}, d['engine']['parity_guard'])"
TITAN_SPEC_FINAL=$sp python scripts/diag_cfg4.py 5 2>&1 | grep -E "^[0-9] estep"; done
