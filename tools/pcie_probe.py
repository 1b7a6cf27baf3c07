"""PCIe probe for the e2e leg: pinned host <-> device copy rates (contiguous, pitched 2-D, 2 streams, duplex)."""
import time, torch
dev = torch.device("cuda:0")
N = 8192 * 8192
h = torch.empty(N, dtype=torch.float32, pin_memory=True); h.zero_()
h2 = torch.empty(N, dtype=torch.float32, pin_memory=True); h2.zero_()
d = torch.empty(N, dtype=torch.float32, device=dev)
d2 = torch.empty(N, dtype=torch.float32, device=dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
gb = N * 4 / 1e9
print(f"H2D contiguous  {gb / t(lambda: d.copy_(h, non_blocking=True)):.1f} GB/s")
print(f"D2H contiguous  {gb / t(lambda: h.copy_(d, non_blocking=True)):.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two_d2h():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print(f"D2H 2 streams   {2 * gb / t(two_d2h):.1f} GB/s aggregate")
def duplex():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
print(f"duplex H2D+D2H  {2 * gb / t(duplex):.1f} GB/s aggregate")
# pitched device (8192 pitch) <-> unpadded host rows of 8191 floats
hp = h[: 8192 * 8191].view(8192, 8191)
dp = d.view(8192, 8192)[:, :8191]
print(f"D2H pitched 2-D {8192 * 8191 * 4 / 1e9 / t(lambda: hp.copy_(dp, non_blocking=True)):.1f} GB/s")
print(f"H2D pitched 2-D {8192 * 8191 * 4 / 1e9 / t(lambda: dp.copy_(hp, non_blocking=True)):.1f} GB/s")
# pageable for reference
pg = torch.empty(N, dtype=torch.float32)
print(f"D2H pageable    {gb / t(lambda: pg.copy_(d)):.1f} GB/s")
