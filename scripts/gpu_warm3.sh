for w in 192 384 512 768 1024; do
echo "Z=3 chunk 1024 warm $w"; python bench.py --steps 10 --warmup 3 --no-em --no-sweep --no-sharded --no-cpu-baseline --chunk 1024 --warm $w | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['engine']['cycles_per_step'], {k: round(v,3) for k,v in d['roofline']['phases_ms'].items()})"
done
for z in 1 2 4; do for w in 192 512; do echo "Z=$z warm $w"; CHUNK=1024 WARM=$w python scripts/diag_cfg4.py $z 1000000 2>&1 | grep -E "^[0-9] estep" ; done; done
