"""Join an ncu SASS source page (ncu -i rep --page source --csv -k kernel) with nvdisasm -gi line info of the same build
and print the warp-stall samples per CUDA source line (innermost inlined line, and the kernel-level line it was inlined at).

  cuobjdump -xelf all mujoco_sim_b200/lib/libb2sim.so; nvdisasm -gi batch.sm_100a.cubin > all.txt
  python tools/dev/src_lines.py all.txt <mangled kernel name> gpurun_out/src/c3_k_smooth_sass.csv [ntop]
"""
import csv, re, sys, collections
dis, kern, page = sys.argv[1], sys.argv[2], sys.argv[3]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 25
# ---- line table of the kernel: offset -> inline chain [(file, line), ...] innermost first ----
table = {}
inside = False
chain, fresh = [], True
pat_file = re.compile(r'//## File "([^"]+)", line (\d+)')
pat_ins = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);')
with open(dis) as f:
    for ln in f:
        if ln.startswith("\t.section\t.text."):
            if inside: break
            inside = (".text." + kern + ",") in ln
            continue
        if not inside: continue
        mf = pat_file.search(ln)
        if mf:
            if fresh: chain = []; fresh = False
            chain.append((mf.group(1).split("/")[-1], int(mf.group(2))))
            continue
        mi = pat_ins.match(ln)
        if mi:
            table[int(mi.group(1), 16)] = list(chain)
            fresh = True
rows = list(csv.reader(open(page)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5 and r[0].startswith("0x")]
for r in data:
    for i in range(2, len(r)):
        if r[i] == "": r[i] = "0"
ci = hdr.index("Instructions Executed"); cs = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][0], 16)
tot_s = sum(int(r[cs]) for r in data); tot_i = sum(int(r[ci]) for r in data)
print(rows[0][1][:90], "| warp inst", tot_i, "samples", tot_s, "| sass", len(data), "line-table", len(table))
inner = collections.defaultdict(lambda: [0, 0, collections.Counter()])
outer = collections.defaultdict(lambda: [0, 0, collections.Counter()])
miss = 0
for r in data:
    off = int(r[0], 16) - base
    ch = table.get(off)
    if not ch: miss += int(r[cs]); continue
    for key, agg in ((ch[0], inner), (ch[-1], outer)):
        a = agg[key]; a[0] += int(r[cs]); a[1] += int(r[ci])
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: a[2][hdr[i][6:]] += v
def show(title, agg):
    print("---- %s (samples%%, inst%%, top stalls) ----" % title)
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
        st = ", ".join("%s %d%%" % (k, 100 * v / max(1, a[0])) for k, v in a[2].most_common(3))
        print("  %5.1f%% %5.1f%%  %s:%d   [%s]" % (100 * a[0] / tot_s, 100 * a[1] / tot_i, key[0], key[1], st))
show("innermost line", inner)
show("kernel-level line", outer)
if miss: print("samples without line info:", miss)
