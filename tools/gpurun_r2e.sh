set -x
for tc in 1 0; do
B2_TC_PROJECT=$tc ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_dense_minv|k_project_tc|k_make_rows|k_make_blocks|k_solve_rows" --launch-skip $((155*4 - 155*tc*0)) -c 8 --csv --log-file gpurun_out/tc${tc}_launches.csv python tools/tick_some.py c4 3 150 > /dev/null 2>&1
done
python - <<'PY'
import csv
for f in ("gpurun_out/tc1_launches.csv", "gpurun_out/tc0_launches.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    print(f)
    for r in rows: print("  ", r[4][:40], r[-4], r[-1], r[-2])
PY
