"""HBM write-only / read-only / copy bandwidth on the GPU box (torch fill_, sum, copy_ over 4 GiB), for the roofline
discussion of the write-dominated kernels (back-projection 12:4 write:read, bilinear up-sampling ~3:1)."""
import torch
n = 1 << 30  # fp32 elements = 4 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, bytes_):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_ / best / 1e6
print("write-only (fill_)      %.0f GB/s" % t(lambda: a.fill_(1.0), 4 * n))
print("write-only (memset 0)   %.0f GB/s" % t(lambda: a.zero_(), 4 * n))
print("read-only  (sum)        %.0f GB/s" % t(lambda: a.sum(), 4 * n))
print("copy (read+write bytes) %.0f GB/s" % t(lambda: b.copy_(a), 8 * n))
