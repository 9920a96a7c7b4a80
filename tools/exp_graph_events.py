"""Experiment: can CUDA events recorded INSIDE a captured graph time individual kernels of a replay?"""
import torch
x = torch.randn(1 << 24, device="cuda")
y = torch.empty_like(x)
evs = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(4)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    g.capture_begin()
    evs[0].record()
    y.copy_(x)
    evs[1].record()
    torch.mul(x, 2.0, out=y)
    evs[2].record()
    for _ in range(10):
        y.add_(1.0)
    evs[3].record()
    g.capture_end()
torch.cuda.synchronize()
for it in range(3):
    g.replay()
    torch.cuda.synchronize()
    print("replay", it, [evs[i].elapsed_time(evs[i + 1]) for i in range(3)])
