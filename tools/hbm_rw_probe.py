"""Developer tool: DRAM bandwidth of this box for a pure write stream, a pure read stream and a copy (torch kernels on
2 GiB buffers, CUDA events, best of 5) -- the write-only figure is the ceiling the trunk's store-heavy junction kernels
run against."""
import json
import torch

n = 1 << 30  # bf16 elements: 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
a.fill_(1.0)
b.fill_(2.0)


def best(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


nbytes = n * 2
res = {
    "write_only_GBs": nbytes / best(lambda: a.zero_()) / 1e6,
    "fill_GBs": nbytes / best(lambda: a.fill_(3.0)) / 1e6,
    "read_only_GBs": nbytes / best(lambda: a.view(torch.int16).max()) / 1e6,
    "copy_read_plus_write_GBs": 2 * nbytes / best(lambda: b.copy_(a)) / 1e6,
    "add_2reads_1write_GBs": 3 * nbytes / best(lambda: torch.add(a, b, out=b)) / 1e6,
}
print(json.dumps(res))
