# Round-2 final measurement pass on the GPU box (one B200): tests, the default bench line, the reference arm, the launch
# list of the bench command, full ncu captures of one steady-state tick (C3, C4, C5; the batch as a single window so that
# every row is a whole-batch kernel), summarised on the box.
set -x
mkdir -p gpurun_out /tmp/rep
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
S=$(date +%s); python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench wall $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/bench_c3.err
python bench.py --impl reference --steps 50 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes); the settle + warm-up ticks precede the timed region
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/launches_c3.log 2>&1
K3='regex:k_pgs_island|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K3" -s 1085 -c 7 -o /tmp/rep/full_c3 python tools/tick_some.py c3 3 150 > /dev/null 2>&1
K4='regex:k_pgs_block|k_make_rows|k_solve_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K4" -s 1200 -c 8 -o /tmp/rep/full_c4 python tools/tick_some.py c4 3 150 > /dev/null 2>&1
B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K3" -s 840 -c 7 -o /tmp/rep/full_c5 python tools/tick_some.py c5 3 120 > /dev/null 2>&1
for c in c3 c4 c5; do python tools/ncu_summary.py /tmp/rep/full_$c.ncu-rep --traffic gpurun_out/traffic_$c.json > gpurun_out/ncu_${c}_summary.txt; done
cuobjdump -sass mujoco_sim_b200/lib/libb2sim.so | grep -E "UTCHMMA|UTCMMA|LDTM|UTCBAR|UTCATOMSWS|UBLKCP|SYNCS|LDGSTS" | awk '{print $2}' | sort | uniq -c > gpurun_out/sass_mnemonics.txt
ls -la gpurun_out /tmp/rep
python tools/bench_line.py gpurun_out/bench_c3.json | cut -c1-400
