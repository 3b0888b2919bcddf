"""HBM bandwidth by access mix on this GPU (torch ops, CUDA events): read-only, write-only, copy, 1:3 read:write."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
N = 1 << 30   # 2 GiB of fp16 per tensor
x = torch.empty(N, dtype=torch.float16, device="cuda").normal_()
y = torch.empty(N, dtype=torch.float16, device="cuda")
q = torch.empty(N // 4, dtype=torch.float16, device="cuda").normal_()
ms = t(lambda: y.fill_(1.0)); print(f"write-only  fill_      {2*N/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
ms = t(lambda: y.zero_()); print(f"write-only  zero_      {2*N/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
ms = t(lambda: x.sum()); print(f"read-only   sum        {2*N/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
ms = t(lambda: y.copy_(x)); print(f"copy 1:1    copy_      {4*N/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
v = y.view(4, N // 4)
ms = t(lambda: v.copy_(q.unsqueeze(0).expand(4, -1))); print(f"read 1 : write 4  bcast {2*(N + N//4)/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
ms = t(lambda: torch.add(x, 1.0, out=y)); print(f"r1:w1 add              {4*N/ms/1e6:8.0f} GB/s  ({ms:.3f} ms)")
