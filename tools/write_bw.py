"""Calibration only: pure-write and copy bandwidth of the device with library kernels (cudaMemset / torch fill / copy)."""
import torch, time
dev = torch.device("cuda")
n = 68763 * 512
Y = torch.empty(n, dtype=torch.float64, device=dev)
Z = torch.empty(n, dtype=torch.float64, device=dev)
def t(fn, reps=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: Y.fill_(1.5)); print("fill  %.1f MB: %.4f ms  %.0f GB/s write" % (n * 8 / 1e6, ms, n * 8 / ms / 1e6))
ms = t(lambda: Y.zero_()); print("zero  %.1f MB: %.4f ms  %.0f GB/s write" % (n * 8 / 1e6, ms, n * 8 / ms / 1e6))
ms = t(lambda: Z.copy_(Y)); print("copy  %.1f MB: %.4f ms  %.0f GB/s read+write" % (n * 8 / 1e6, ms, 2 * n * 8 / ms / 1e6))
big = torch.empty(1 << 28, dtype=torch.float64, device=dev)
ms = t(lambda: big.fill_(1.5), 10); print("fill  %.1f MB: %.4f ms  %.0f GB/s write" % (big.numel() * 8 / 1e6, ms, big.numel() * 8 / ms / 1e6))
