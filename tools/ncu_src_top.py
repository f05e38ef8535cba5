"""Developer tool: top stall-sample SASS lines of an ncu report (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
which = int(sys.argv[3]) if len(sys.argv) > 3 else len(starts) - 1   # kernel index inside the report
start = starts[which]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(lines)
rows = list(csv.DictReader(lines[start:end]))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot)
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    st = {c[6:]: int(r[c]) for c in stall_cols if int(r[c] or 0)}
    print(f"{i:5d} {int(r['# Samples']):6d} {r['Source'][:90]:90s} {st}")
