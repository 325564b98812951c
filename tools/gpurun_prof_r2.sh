# Round-2 measurement pass on the GPU box: tests, the default bench line, the reference arm, launch lists, full ncu
# captures of one steady-state tick (C3, C4, C4 with the tensor-core projection), summarised on the box.
set -x
mkdir -p gpurun_out /tmp/rep
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 400 gpurun_out/bench_c3.err
python bench.py --impl reference --steps 50 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes); 160 ticks x 9 kernels precede the timed region
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/launches_c3.log 2>&1
K3='regex:k_pgs_island|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
ncu --set full --clock-control none --import-source on -k "$K3" -s 1085 -c 7 -o /tmp/rep/full_c3 python tools/tick_some.py c3 3 150 > /dev/null 2>&1
K4='regex:k_pgs_block|k_make_rows|k_solve_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
ncu --set full --clock-control none --import-source on -k "$K4" -s 1240 -c 8 -o /tmp/rep/full_c4 python tools/tick_some.py c4 3 150 > /dev/null 2>&1
KT='regex:k_project_tc|k_dense_minv'
B2_TC_PROJECT=1 ncu --set full --clock-control none --import-source on -k "$KT" -s 310 -c 2 -o /tmp/rep/full_c4tc python tools/tick_some.py c4 3 150 > /dev/null 2>&1
for c in c3 c4 c4tc; do python tools/ncu_summary.py /tmp/rep/full_$c.ncu-rep --traffic gpurun_out/traffic_$c.json > gpurun_out/ncu_${c}_summary.txt; done
ls -la gpurun_out /tmp/rep
python tools/exp_tc.py > gpurun_out/tc_vs_ffma.txt 2>&1
tail -1 gpurun_out/bench_c3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c3', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],3), 'ms e2e', round(d['e2e']['value']/1e6,3), {x: round(v,3) for x,v in r['kernel_ms_all'].items() if v>0}, 'hbm', round(r['frac'],5), 'fp32', round(r['fp32_frac'],5))
print('cpu', d.get('cpu_baseline')); print('drift', d.get('drift')); print('exch', d.get('obs_exchange'))
for c,v in d.get('configs',{}).items(): print(c, {k: (round(x,4) if isinstance(x,float) else x) for k,x in v.items() if k not in ('workload','kernel_ms_all')})
"
