"""Developer tool: summarise `ncu --page source --csv --print-source sass` of one kernel: executed warp-instructions and
stall samples by opcode class, and the hottest instructions.  usage: ncu_source_summary.py file.csv [points]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
# a report with several launches holds one section per launch ("Kernel Name" line + header + instructions): -k N picks the N-th
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if "-k" in sys.argv and starts:
    k = int(sys.argv[sys.argv.index("-k") + 1])
    rows = rows[starts[k]:(starts[k + 1] if k + 1 < len(starts) else len(rows))]
    print("launch", k, ":", rows[0][1][:100])
elif len(starts) > 1:
    rows = rows[starts[0]:starts[1]]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
npts = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else None
tot_inst = 0; tot_samp = 0
by_op = collections.Counter(); samp_op = collections.Counter(); stall_tot = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    n = float(r[col["Instructions Executed"]] or 0); s = float(r[col["# Samples"]] or 0)
    tot_inst += n; tot_samp += s
    by_op[op] += n; samp_op[op] += s
    for h in stall_cols:
        stall_tot[h] += float(r[col[h]] or 0)
    recs.append((s, n, r[col["Address"]], src, {h: float(r[col[h]] or 0) for h in stall_cols}))
scale = 32.0 / npts if npts else 1.0
print("total warp-instructions %.4g  (per point: %.1f)   samples %d" % (tot_inst, tot_inst * scale, tot_samp))
print("%-10s %12s %8s %8s" % ("opcode", "inst/pt" if npts else "inst", "inst%", "samp%"))
for op, n in by_op.most_common(28):
    print("%-10s %12.2f %7.1f%% %7.1f%%" % (op, n * scale, 100 * n / tot_inst, 100 * samp_op[op] / max(tot_samp, 1)))
print("stall totals:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / max(tot_samp, 1)) for h, v in stall_tot.most_common(9)))
if "-top" in sys.argv:
    for s, n, addr, src, st in sorted(recs, key=lambda x: -x[0])[:40]:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("%6d samp %10.3g inst  %s  %-60s %s" % (s, n, addr, src[:60], " ".join("%s:%d" % (k[6:], v) for k, v in top)))
