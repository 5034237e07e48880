import torch
n = 4 * 1024**3
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, name, bytes_):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  {bytes_ / ms / 1e9:.2f} TB/s")
t(lambda: a.zero_(), "write-only (memset 4 GiB)", n)
t(lambda: a.fill_(3), "write-only (fill kernel)", n)
t(lambda: b.copy_(a), "copy (read+write)", 2 * n)
af = a.view(torch.float32)
t(lambda: af.sum(), "read-only (sum)", n)
