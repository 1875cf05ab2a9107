python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "estep or em_ or config4 or chunked or full_size or edge or bitwise or kernel_variants or legacy or dense or generic" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-em --no-sweep --no-sharded --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['engine']['cycles_per_step'], {k: round(v,3) for k,v in d['roofline']['phases_ms'].items()}, d['engine']['parity_guard'])"
python scripts/diag_cfg4.py 5 2>&1 | grep -E "^[0-9] estep"
