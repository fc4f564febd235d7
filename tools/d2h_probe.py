"""Device -> host options for a one-off 120 MB chain (c2): pinned staging vs pageable."""
import time, numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
d1 = torch.randn((5000, 1000, 2), dtype=torch.float64, device="cuda"); d2 = torch.randn((5000, 1000), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
def T(f, label):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{label}: {(time.perf_counter() - t) * 1e3:.1f} ms"); return r
h = T(lambda: (torch.empty(d1.shape, dtype=d1.dtype, pin_memory=True), torch.empty(d2.shape, dtype=d2.dtype, pin_memory=True)), "pinned alloc 80+40 MB (torch caching host allocator)")
T(lambda: (h[0].copy_(d1, non_blocking=True), h[1].copy_(d2, non_blocking=True)), "D2H into pinned")
del h
p = T(lambda: (torch.empty(d1.shape, dtype=d1.dtype), torch.empty(d2.shape, dtype=d2.dtype)), "pageable alloc")
T(lambda: (p[0].copy_(d1), p[1].copy_(d2)), "D2H into fresh pageable (first touch inside the copy)")
T(lambda: (p[0].copy_(d1), p[1].copy_(d2)), "D2H into touched pageable")
q = T(lambda: (torch.zeros(d1.shape, dtype=d1.dtype), torch.zeros(d2.shape, dtype=d2.dtype)), "pageable alloc + zero fill")
T(lambda: (q[0].copy_(d1), q[1].copy_(d2)), "D2H into zero-filled pageable")
n = T(lambda: (np.empty(d1.shape), np.empty(d2.shape)), "numpy empty")
T(lambda: (torch.from_numpy(n[0]).copy_(d1), torch.from_numpy(n[1]).copy_(d2)), "D2H into numpy empty")
cudart = torch.cuda.cudart()
r = T(lambda: (torch.empty(d1.shape, dtype=d1.dtype), torch.empty(d2.shape, dtype=d2.dtype)), "pageable alloc")
T(lambda: [cudart.cudaHostRegister(x.data_ptr(), x.numel() * 8, 0) for x in r], "cudaHostRegister 120 MB exact size")
T(lambda: (r[0].copy_(d1, non_blocking=True), r[1].copy_(d2, non_blocking=True)), "D2H into registered")
T(lambda: [cudart.cudaHostUnregister(x.data_ptr()) for x in r], "unregister")
