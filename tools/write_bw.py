"""Pure-write and copy bandwidth of this GPU (torch fill_ / copy_ on 1 GiB): the ceiling for write-dominated kernels."""
import torch
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
def t(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
print("fill_  GB/s", x.numel() * 4 / t(lambda: x.fill_(1.0)) / 1e9)
print("zero_  GB/s", x.numel() * 4 / t(lambda: x.zero_()) / 1e9)
print("copy_  GB/s (read + write)", 2 * x.numel() * 4 / t(lambda: y.copy_(x)) / 1e9)
