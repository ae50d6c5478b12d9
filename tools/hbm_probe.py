import torch, time
x = torch.empty(1310720000 // 2, dtype=torch.float16, device="cuda")
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
ms = t(lambda: x.zero_()); print("fill 1.31GB ms", ms, "TB/s", 1.31072/ms)
ms = t(lambda: y.copy_(x)); print("copy 1.31GB ms", ms, "TB/s r+w", 2*1.31072/ms)
ms = t(lambda: x.sum()); print("read 1.31GB ms", ms, "TB/s", 1.31072/ms)
