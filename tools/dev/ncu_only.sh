mkdir -p gpurun_out /tmp/rep
K3='regex:k_pgs_island|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
B2_NO_ORDER_FORK=1 B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K3" -s 1085 -c 7 -o /tmp/rep/full_c3 python tools/tick_some.py c3 3 150 > /dev/null 2>&1
K4='regex:k_pgs_block|k_make_rows|k_solve_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_order_envs'
B2_NO_ORDER_FORK=1 B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K4" -s 1200 -c 8 -o /tmp/rep/full_c4 python tools/tick_some.py c4 3 150 > /dev/null 2>&1
B2_NO_ORDER_FORK=1 B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K3" -s 840 -c 7 -o /tmp/rep/full_c5 python tools/tick_some.py c5 3 120 > /dev/null 2>&1
for c in c3 c4 c5; do python tools/ncu_summary.py /tmp/rep/full_$c.ncu-rep --traffic gpurun_out/traffic_$c.json > gpurun_out/ncu_${c}_summary.txt; done
wc -c gpurun_out/ncu_c?_summary.txt; cat gpurun_out/traffic_c3.json
