"""Write-only / copy bandwidth of this GPU at the size of the cfg2 output (4.3 GB)."""
import torch
n = 2 ** 28
x = torch.empty(n, dtype=torch.complex128, device="cuda")
y = torch.empty(n, dtype=torch.complex128, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.zero_()); print("memset 4.3 GB: %.3f ms  %.0f GB/s (write only)" % (ms, n * 16 / ms / 1e6))
ms = t(lambda: x.fill_(1.5)); print("fill   4.3 GB: %.3f ms  %.0f GB/s (write only)" % (ms, n * 16 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy   4.3 GB: %.3f ms  %.0f GB/s (read + write)" % (ms, 2 * n * 16 / ms / 1e6))
