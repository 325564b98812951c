mkdir -p /tmp/rep gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pgs_island -s 155 -c 1 -o /tmp/rep/pgs python tools/tick_some.py c3 3 150 > /dev/null 2>&1
ncu -i /tmp/rep/pgs.ncu-rep --page source --csv > /tmp/rep/pgs_source.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("/tmp/rep/pgs_source.csv")))
hdr = rows[0]
print(hdr[:12])
# find columns
def col(name):
    for i, h in enumerate(hdr):
        if h.strip() == name: return i
    return None
cs = col("Source"); ci = col("# Inst Executed") or col("Instructions Executed"); cw = col("Warp Stall Sampling (All Samples)") or col("# Samples")
print("cols", cs, ci, cw)
for i, h in enumerate(hdr): print(i, h)
PY
head -c 3000 /tmp/rep/pgs_source.csv > gpurun_out/pgs_source_head.txt
cp /tmp/rep/pgs_source.csv gpurun_out/pgs_source.csv
ls -la gpurun_out/pgs_source.csv
