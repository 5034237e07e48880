"""HBM bandwidth probes on one B200: write-only, read-only and copy streams with the store/load
patterns the product kernels use.  CUDA-event times here; run the same script under
`ncu --metrics dram__bytes_write.sum.per_second,dram__bytes_read.sum.per_second,gpu__time_duration.sum` for the DRAM
counters.  Output is committed as profiles/r2_bw_probe.txt."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L

GiB = 1 << 30
n = 4 * GiB
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
lib = L.lib()
reps = 1 if "--once" in sys.argv else 5


def run(mode, ctas, depth=1, src=None):
    L.call("nmx_diag_bw", L.i32(mode), L.ptr(a), L.ptr(src), L.i64(n), L.i32(ctas), L.i32(depth), L.stream())


def t(name, fn, bytes_):
    for _ in range(0 if reps == 1 else 2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:78s} {ms:8.3f} ms  {bytes_ / ms / 1e6:8.1f} GB/s", flush=True)


print(f"buffer {n / GiB:.0f} GiB, {reps} reps (CUDA events)")
t("write-only: cudaMemsetAsync", lambda: run(0, 1), n)
t("write-only: torch fill_ (vectorised elementwise kernel)", lambda: a.fill_(3), n)
for ctas in (148 * 4, 148 * 8, 148 * 16):
    t(f"write-only: st.global.v4 grid-stride fill, {ctas} CTAs x 256 thr", lambda c=ctas: run(1, c), n)
for depth in (1, 2, 3):
    t(f"write-only: bulk-async 1-D stores smem->global, 148 CTAs, 64 KB tiles, {depth} group(s) in flight",
      lambda d=depth: run(2, 148, d), n)
for depth in (1, 2, 3):
    t(f"write-only: TMA 2-D stores, chain pattern (128x64 boxes of [P,256] bf16), 148 CTAs, {depth} in flight",
      lambda d=depth: run(5, 148, d), n)
t("read-only: ld.global.v4 grid-stride, 2368 CTAs", lambda: run(3, 148 * 16), n)
t("read-only: TMA 2-D tile loads (128x64 boxes), 148 CTAs, 2 x 64 KB in flight", lambda: run(6, 148), n)
t("read-only: torch sum (fp32)", lambda: a.view(torch.float32).sum(), n)
t("copy: ld/st.global.v4 grid-stride, 2368 CTAs (read + write bytes)", lambda: run(4, 148 * 16, 1, b), 2 * n)
t("copy: torch copy_ (read + write bytes; MEASURED_PEAKS.json method)", lambda: a.copy_(b), 2 * n)
