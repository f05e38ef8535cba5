"""Developer tool: writes profiles/r02_traffic.json (DRAM bytes per launch group, read by bench.py for the `traffic`
fields) from the committed ncu passes: profiles/r02_ncu_trunk_traffic.csv (one trunk call at batch 256) and
profiles/r02_ncu_pre_launches.csv (preprocess calls at batch 256)."""
import collections, csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(path):
    rows = collections.OrderedDict()
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(r['ID'], {'name': r['Kernel Name']})
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r['Metric Unit'], 1)
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * scale
    return list(rows.values())
out = {}
t = load(os.path.join(ROOT, "profiles", "r02_ncu_trunk_traffic.csv"))
s = next(i for i, d in enumerate(t) if 'stem_pool' in d['name'])
t = (t[s:] + t[:s])[:45]
out["trunk_call_batch256"] = {"bytes": sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in t),
                              "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 45 kernels of one trunk "
                                        "call at batch 256, profiles/r02_ncu_trunk_traffic.csv"}
p = load(os.path.join(ROOT, "profiles", "r02_ncu_pre_launches.csv"))
calls = [d for d in p if 'resample_fused' in d['name']]
plans = [d for d in p if 'resample_plan' in d['name']]
n = min(len(calls), len(plans))
out["preprocess_call_batch256"] = {
    "bytes": sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in calls[:n] + plans[:n]) / max(n, 1),
    "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of resample_plan_kernel + resample_fused_kernel, mean of {n} "
              "calls at batch 256 of the bench workload, profiles/r02_ncu_pre_launches.csv"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
