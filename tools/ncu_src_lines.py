"""Developer tool: stall samples and executed instructions of an ncu report aggregated per CUDA source line."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
hdr_idx = [i for i, l in enumerate(lines) if l.startswith('"Line No","Source","Address"')]
agg = collections.OrderedDict()
for hi, h in enumerate(hdr_idx):
    end = hdr_idx[hi + 1] if hi + 1 < len(hdr_idx) else len(lines)
    rd = csv.reader(lines[h:end]); header = next(rd)
    iS = header.index("# Samples"); iI = header.index("Instructions Executed")
    cur = None
    for row in rd:
        if len(row) < len(header): continue
        if row[0]: cur = (int(row[0]), row[1].strip()[:100])
        if cur is None or not row[2]: continue
        a = agg.setdefault(cur, [0, 0])
        try: a[0] += int(row[iS] or 0); a[1] += int(row[iI] or 0)
        except ValueError: pass
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print("total samples", tot, "instructions", toti)
for (ln, src), (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {100*s_/max(tot,1):5.1f}% samp {100*i_/max(toti,1):5.1f}% inst  {src}")
