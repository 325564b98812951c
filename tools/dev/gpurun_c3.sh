python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in c3 c4 c5; do
python bench.py --config $c --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_all']; print('$c', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],3), {x: round(v,3) for x,v in k.items() if v>0})"
done
