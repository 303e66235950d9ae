"""PCIe ceiling of the e2e leg: pinned H2D, D2H and simultaneous both-ways bandwidth for buffers of the
e2e step's size (206 MB in, 208 MB out), CUDA-event timed.    python tools/pcie_probe.py"""
import torch

dev = torch.device("cuda", 0)
n = 206 * 1000 * 1000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 4 / 1e9
for name, fn, total in (("H2D", h2d, gb), ("D2H", d2h, gb), ("both ways at once", both, 2 * gb)):
    ms = timed(fn)
    print(f"{name:18s} {ms:6.2f} ms per {total * 1e3:.0f} MB  -> {total / ms * 1e3:6.1f} GB/s")
