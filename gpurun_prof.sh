set -x
mkdir -p gpurun_out
# launch lists (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 120 --csv --log-file gpurun_out/launches_c3.csv python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_c3.log 2>&1
# full captures of the kernels of one steady-state tick
ncu --set full --clock-control none --import-source on -k regex:k_chain_team -s 4 -c 1 -o gpurun_out/full_c2 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_pgs_block|k_make_constraint|k_smooth|k_collide|k_integrate|k_order_envs' -s 930 -c 6 -o gpurun_out/full_c3 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
