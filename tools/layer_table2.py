"""Developer tool: per-launch table of one trunk call from the csv of tools/ncu_launches.sh."""
import csv, sys, collections
path = sys.argv[1]
rows = collections.OrderedDict()
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    d = rows.setdefault(r['ID'], {'name': r['Kernel Name']})
    d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    d['unit_' + r['Metric Name']] = r['Metric Unit']
tot = 0; tw = 0
for i, (k, d) in enumerate(rows.items()):
    us = d['gpu__time_duration.sum']
    if d['unit_gpu__time_duration.sum'] in ('nsecond', 'ns'): us /= 1e3
    def mb(key):
        v = d.get(key, 0); u = d.get('unit_' + key, 'byte')
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(u, 1e-6)
    rd, wr = mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum')
    tp = d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)
    name = d['name'].replace('void ', '').replace('(ConvParams)', '').replace('(ChainParams)', '')[:44]
    print(f"{i:2d} {name:44s} {us:7.1f} us  {rd+wr:7.0f} MB  {(rd+wr)/us/1e0*1e0/1e0:7.0f} MB/us*  tensor {tp:5.1f}%".replace('MB/us*', 'GB/s x1e-3'))
    tot += us; tw += us * tp
print(f"total {tot:.0f} us, tensor pipe time-weighted {tw/tot:.1f}%")
