import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
groups = collections.defaultdict(lambda: [0, 0, 0, []])
tot_s = 0; tot_i = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    try: ex = int(r[iex] or 0); sm = int(r[ism] or 0)
    except Exception: continue
    g = groups[ex]; g[0] += 1; g[1] += sm; g[2] += ex; tot_s += sm; tot_i += ex
    if sm > 0: g[3].append((sm, r[isrc].strip()[:70]))
print("total samples", tot_s, "total warp-inst", tot_i)
for ex, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:12]:
    print("exec/instr=%10d  n_sass=%4d  samples=%6d (%4.1f%%)  inst=%5.1f%%" % (ex, g[0], g[1], 100.0 * g[1] / tot_s, 100.0 * g[2] / tot_i))
    for sm, src in sorted(g[3], reverse=True)[:6]:
        print("        %6d  %s" % (sm, src))
