import csv, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r: hdr = r; start = i + 1; break
print("cols:", [h for h in hdr][:8])
isrc = hdr.index("Source"); iex = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
iln = hdr.index("#") if "#" in hdr else None
data = []
for r in rows[start:]:
    if len(r) <= iex: continue
    try: data.append((int(r[iex] or 0), int(r[ism] or 0), r[0], r[isrc]))
    except Exception: pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total inst", tot, "samples", ts)
for d in sorted(data, key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%5s inst %5.1f%% samp %5.1f%%  %s" % (d[2], 100.0 * d[0] / max(tot,1), 100.0 * d[1] / max(ts,1), d[3].strip()[:110]))
