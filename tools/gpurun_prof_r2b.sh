set -x
mkdir -p gpurun_out /tmp/rep
K4='regex:k_pgs_block|k_make_rows|k_solve_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
ncu --set full --clock-control none --import-source on -k "$K4" -s 1200 -c 8 -o /tmp/rep/full_c4 python tools/tick_some.py c4 3 150 > /dev/null 2>&1
KT='regex:k_project_tc|k_dense_minv'
B2_TC_PROJECT=1 ncu --set full --clock-control none --import-source on -k "$KT" -s 300 -c 2 -o /tmp/rep/full_c4tc python tools/tick_some.py c4 3 150 > /dev/null 2>&1
K5='regex:k_pgs_island|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
ncu --set full --clock-control none --import-source on -k "$K5" -s 840 -c 7 -o /tmp/rep/full_c5 python tools/tick_some.py c5 3 120 > /dev/null 2>&1
for c in c4 c4tc c5; do python tools/ncu_summary.py /tmp/rep/full_$c.ncu-rep --traffic gpurun_out/traffic_$c.json > gpurun_out/ncu_${c}_summary.txt; done
ls -la /tmp/rep gpurun_out/ncu_*
cuobjdump -sass mujoco_sim_b200/lib/libb2sim.so | grep -E "UTCHMMA|UTCMMA|LDTM|UTCBAR|UTCATOMSWS|UBLKCP|SYNCS" | awk '{print $2}' | sort | uniq -c > gpurun_out/sass_mnemonics.txt
cat gpurun_out/sass_mnemonics.txt
