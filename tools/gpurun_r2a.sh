set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.err
tail -1 gpurun_out/bench_default.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('c3', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],3), 'ms rep', d['repeats'], 'e2e', round(d['e2e']['value']/1e6,3), {x: round(v,3) for x,v in r['kernel_ms_all'].items() if v>0}, 'hbm', r['frac'], 'fp32', r['fp32_frac'])
print('cpu', d.get('cpu_baseline'))
print('drift', d.get('drift'))
for c,v in d.get('configs',{}).items(): print(c, v)
print('clocks', d['clocks'])
"
