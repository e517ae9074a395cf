"""Pinned-memory copy bandwidth of this box: H2D alone, D2H alone, both at once (tools only)."""
import time
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t) / 1e9


run(True, True, 2)
print("H2D alone GB/s", run(True, False))
print("D2H alone GB/s", run(False, True))
print("both, per direction GB/s", run(True, True))
