"""Developer tool: per-launch markdown table (layer, kernel, us, TFLOP/s, DRAM MB and GB/s, tensor-pipe %) of ONE trunk
call from the csv of tools/ncu_launches.sh, and the DRAM-traffic entry of profiles/r02_traffic.json.
usage: python tools/layer_table_md.py gpurun_out/r02_ncu_trunk_traffic.csv > profiles/r02_trunk_layer_table.md"""
import collections, csv, sys
path = sys.argv[1]; B = 256
rows = collections.OrderedDict()
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    d = rows.setdefault(r['ID'], {'name': r['Kernel Name']})
    d[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])
rows = list(rows.values())
start = next(i for i, d in enumerate(rows) if 'stem_pool' in d['name'])
rows = (rows[start:] + rows[:start])[:45]
# launch order of resnet50_forward (csrc/conv.cu)
labels = ["stem 7x7/2 + max pool"]
planes, blocks, inpl, hw = [64, 128, 256, 512], [3, 4, 6, 3], 64, 56
flops = [2 * B * 112 * 112 * 64 * 147]
conv1_done = False
for l in range(4):
    for b in range(blocks[l]):
        s = 2 if (b == 0 and l > 0) else 1
        w = planes[l]; ho = hw // s; name = f"L{l+1}.{b}"
        has_ds = b == 0
        nxt_same_res = not (b == blocks[l] - 1 and l == 3)
        chain = (w * 4 <= 512) and nxt_same_res
        f1 = 2 * B * hw * hw * w * inpl
        if not conv1_done:
            labels.append(f"{name} conv1 {inpl}->{w} @{hw}"); flops.append(f1)
        labels.append(f"{name} conv2 3x3{'/2' if s == 2 else ''} {w}->{w} @{ho}"); flops.append(2 * B * ho * ho * w * w * 9)
        fds = 2 * B * ho * ho * w * 4 * inpl
        folded = has_ds and s == 1 and chain
        if has_ds and not folded:
            labels.append(f"{name} downsample {inpl}->{w*4}{'/2' if s == 2 else ''}"); flops.append(fds)
        f3 = 2 * B * ho * ho * w * 4 * w
        if chain:
            nw = planes[l] if b < blocks[l] - 1 else planes[l + 1]
            fn = 2 * B * ho * ho * nw * w * 4
            labels.append(f"{name} conv3{' + shortcut conv' if folded else ' + res'} + next conv1 ({w*4}->{nw}) @{ho}")
            flops.append(f3 + fn + (fds if folded else 0)); conv1_done = True
        else:
            last = l == 3 and b == blocks[l] - 1
            labels.append(f"{name} conv3 {w}->{w*4} + res{' + global average pool' if last else ''} @{ho}")
            flops.append(f3); conv1_done = False
        inpl = w * 4; hw = ho
assert len(labels) == 45, len(labels)
def val(d, key, scale):
    v, u = d.get(key, (0, ''))
    return v * scale.get(u, 1)
print("# Trunk per-launch table, batch 256 (round 2)\n")
print(f"From `{path.split('/')[-1]}` (ncu `--clock-control none`: cold cache, serialised launches; the 45 launches of ONE")
print("`irp_resnet50_embed` call). FLOP = true conv FLOPs of the launch.\n")
print("| # | launch | kernel | µs | TFLOP/s | DRAM GB/s | DRAM MB | tensor pipe % |\n|---|---|---|---|---|---|---|---|")
tot = tot_f = tot_b = tw = 0
for i, (d, lab, fl) in enumerate(zip(rows, labels, flops)):
    us = val(d, 'gpu__time_duration.sum', {'nsecond': 1e-3, 'ns': 1e-3, 'usecond': 1, 'us': 1})
    mb = val(d, 'dram__bytes_read.sum', {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}) + \
         val(d, 'dram__bytes_write.sum', {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3})
    tp = d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', (0, ''))[0]
    kn = d['name'].replace('void ', '').split('(')[0]
    print(f"| {i} | {lab} | `{kn}` | {us:.1f} | {fl/us/1e6:.0f} | {mb/us*1e3:.0f} | {mb:.0f} | {tp:.1f} |")
    tot += us; tot_f += fl; tot_b += mb; tw += us * tp
print(f"| | **one call** | 45 launches | **{tot:.0f}** | **{tot_f/tot/1e6:.0f}** | {tot_b/tot*1e3:.0f} | **{tot_b:.0f}** | {tw/tot:.1f} (time-weighted) |")
sys.stderr.write(f"trunk_call_batch256 bytes {tot_b*1e6:.0f}\n")
