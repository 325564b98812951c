set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --config c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json
python bench.py --config c2 --nenv 262144 --no-cpu-baseline > gpurun_out/bench_c2_256k.json 2> gpurun_out/bench_c2_256k.err
python bench.py --config c3 --steps 20 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --config c4 --steps 10 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --config c5 --steps 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python bench.py --config c2 --impl reference > gpurun_out/bench_ref_c2.json 2>&1
# launch lists (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 60 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c4.log 2>&1
# full captures of the kernels of one steady-state tick; summarised on the box (the reports exceed the 64 MiB return limit)
mkdir -p /tmp/rep
ncu --set full --clock-control none --import-source on -k regex:k_chain_team -s 4 -c 1 -o /tmp/rep/full_c2 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_pgs_block|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs' -s 1085 -c 7 -o /tmp/rep/full_c3 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_pgs_block|k_make_rows|k_solve_rows|k_make_blocks|k_smooth|k_collide|k_integrate' -s 1085 -c 7 -o /tmp/rep/full_c4 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for c in c2 c3 c4; do python tools/ncu_summary.py /tmp/rep/full_$c.ncu-rep --traffic gpurun_out/traffic_$c.json > gpurun_out/ncu_${c}_summary.txt; done
ls -la gpurun_out /tmp/rep
for c in c3 c4 c5; do
tail -1 gpurun_out/bench_$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_all']; print('$c', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,3), {x: round(v,3) for x,v in k.items() if v>0}, d.get('cpu_baseline',{}).get('value'), d.get('drift'))"
done
