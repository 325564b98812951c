# source-level stall sampling of one steady-state C3 tick: one report, then per-kernel CUDA-line and SASS views
mkdir -p /tmp/rep gpurun_out/src
CFG=${1:-c3}
K='regex:k_pgs_island|k_pgs_block|k_make_rows|k_make_blocks|k_smooth|k_collide|k_integrate|k_solve_rows'
B2_SUBBATCH=1 ncu --set full --clock-control none --import-source on -k "$K" -s ${2:-900} -c ${3:-6} -o /tmp/rep/src_$CFG python tools/tick_some.py $CFG 3 150 > gpurun_out/src/ncu_$CFG.log 2>&1; tail -5 gpurun_out/src/ncu_$CFG.log
for k in k_pgs_island k_pgs_block k_make_rows k_make_blocks k_smooth k_collide k_integrate k_solve_rows; do
  ncu -i /tmp/rep/src_$CFG.ncu-rep --page source --csv --print-source cuda -k regex:$k > gpurun_out/src/${CFG}_${k}_cuda.csv 2>/dev/null
  ncu -i /tmp/rep/src_$CFG.ncu-rep --page source --csv -k regex:$k > gpurun_out/src/${CFG}_${k}_sass.csv 2>/dev/null
done
python tools/ncu_summary.py /tmp/rep/src_$CFG.ncu-rep > gpurun_out/src/${CFG}_summary.txt
ls -la gpurun_out/src
