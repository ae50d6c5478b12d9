"""Does replaying one sampling call from a CUDA graph (no launch gaps) beat the eager launch sequence?  Timing experiment."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
sys.argv = ["bench"]
import bench


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


gd, cfg, _ = bench.build_model(torch.device("cuda"))
for B, K in ((64, 4), (8, 4), (1, 4)):
    cond = torch.randn(B, 1000, 256, device="cuda")
    xT = torch.rand(B, 1, 128, 1000, device="cuda")
    with torch.no_grad():
        eager = timed(lambda: gd(cond, infer_step=K, infer=True, seed=1, x_T=xT))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            gd(cond, infer_step=K, infer=True, seed=1, x_T=xT)
            with torch.cuda.graph(g, stream=s):
                out = gd(cond, infer_step=K, infer=True, seed=1, x_T=xT)
        torch.cuda.current_stream().wait_stream(s)
        graph = timed(g.replay)
    print(f"B={B} K={K}: eager {eager:.3f} ms, graph replay {graph:.3f} ms ({100 * (eager - graph) / eager:.1f}% faster)", flush=True)
