import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ci = hdr.index("Instructions Executed"); cs = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ci]) for r in data); tot_samp = sum(int(r[cs]) for r in data)
print(rows[0][1][:80], "| inst", tot_inst, "samples", tot_samp)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print("  stalls:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, tot_samp)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]))
mix, smp = {}, {}
for r in data:
    t = r[1].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[ci]); smp[op] = smp.get(op, 0) + int(r[cs])
print("  mix:", ", ".join("%s %.0f%%/%.0f%%" % (k, 100 * v / tot_inst, 100 * smp[k] / tot_samp) for k, v in sorted(mix.items(), key=lambda x: -smp[x[0]])[:10]))
for r in sorted(data, key=lambda r: -int(r[cs]))[:14]:
    print("   %6s smp %9s inst  %s" % (r[cs], r[ci], r[1].strip()[:100]))
