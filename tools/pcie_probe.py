"""Pinned device->host bandwidth of this box: one stream, two streams, and a pitched 2-D copy shape (tools only)."""
import torch, time
n = 300 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def one():
    with torch.cuda.stream(s1): h1.copy_(d, non_blocking=True)
def two():
    with torch.cuda.stream(s1): h1[: n // 2].copy_(d[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h1[n // 2 :].copy_(d[n // 2 :], non_blocking=True)
def both():
    with torch.cuda.stream(s1): h1.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d, non_blocking=True)
def up():
    with torch.cuda.stream(s1): d.copy_(h1, non_blocking=True)
def updown():
    with torch.cuda.stream(s1): d.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d, non_blocking=True)
print("D2H one stream      %.1f GB/s" % (n / run(one) / 1e9))
print("D2H split on two    %.1f GB/s" % (n / run(two) / 1e9))
print("D2H two full copies %.1f GB/s (sum)" % (2 * n / run(both) / 1e9))
print("H2D one stream      %.1f GB/s" % (n / run(up) / 1e9))
print("H2D + D2H together  %.1f GB/s (sum)" % (2 * n / run(updown) / 1e9))
# pitched rows like the grid mirrors: 4097 doubles out of a 4160-double pitch
rows, w, pitch = 4096, 4097 * 8, 4160 * 8
dd = torch.empty(rows * pitch, dtype=torch.uint8, device="cuda").view(rows, pitch)
hh = torch.empty(rows * w, dtype=torch.uint8).pin_memory().view(rows, w)
def pitched():
    with torch.cuda.stream(s1): hh.copy_(dd[:, :w], non_blocking=True)
print("D2H pitched 2-D     %.1f GB/s" % (rows * w / run(pitched) / 1e9))
