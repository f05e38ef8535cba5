"""Developer tool: per-kernel stall-reason totals and the top CUDA source lines (samples, instructions, top stall
reasons) of an `ncu --page source --csv --print-source cuda,sass` dump.  usage: ncu_src_stalls.py dump.csv [top] [kernel indices]"""
import csv, collections, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
which = [int(x) for x in sys.argv[3:]]
lines = open(path).read().splitlines()
hdr_idx = [i for i, l in enumerate(lines) if l.startswith('"Line No","Source","Address"')]
print(len(hdr_idx), 'kernels')
for hi, h in enumerate(hdr_idx):
    if which and hi not in which: continue
    end = hdr_idx[hi + 1] if hi + 1 < len(hdr_idx) else len(lines)
    rd = csv.reader(lines[h:end]); header = next(rd)
    iS = header.index('# Samples'); iI = header.index('Instructions Executed')
    stall_cols = [i for i, c in enumerate(header) if c.startswith('stall_')]
    agg = collections.OrderedDict(); cur = None; st = collections.Counter()
    for row in rd:
        if len(row) < len(header): continue
        if row[0]: cur = (int(row[0]), row[1].strip()[:90])
        if cur is None or not row[2]: continue
        a = agg.setdefault(cur, [0, 0, collections.Counter()])
        try:
            a[0] += int(row[iS] or 0); a[1] += int(row[iI] or 0)
            for i in stall_cols:
                v = int(row[i] or 0); a[2][header[i]] += v; st[header[i]] += v
        except ValueError: pass
    tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
    print('=== kernel', hi, 'samples', tot, 'inst', toti)
    print(st.most_common(8))
    for (ln, src), (s_, i_, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{ln:5d} {100*s_/max(tot,1):5.1f}% samp {100*i_/max(toti,1):5.1f}% inst {src} | {c.most_common(2)}")
