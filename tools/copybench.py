#!/usr/bin/env python
"""PCIe bandwidth of the SM-driven copies (pc_copy_async) next to the copy engines (development tool)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding  # noqa: E402

ctx = binding.Context(0)
n = 512 << 20
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.random_(0, 255)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, pin_memory=True)
side = torch.cuda.Stream()                      # pc_copy_async(stream = 0) means "the context stream": use a real handle
st = side.cuda_stream


def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n / best / 1e6


print("copy engine H2D %.1f GB/s  D2H %.1f GB/s" % (timeit(lambda: dev.copy_(host, non_blocking=True)),
                                                    timeit(lambda: back.copy_(dev, non_blocking=True))))
for ctas in (4, 8, 16, 24, 32, 64, 148):
    h2d = timeit(lambda: ctx.copy_async(dev, host, stream=st, n_ctas=ctas))
    d2h = timeit(lambda: ctx.copy_async(back, dev, stream=st, n_ctas=ctas))
    print("SM copy %3d CTAs  H2D %.1f GB/s  D2H %.1f GB/s  ok=%s" % (ctas, h2d, d2h, bool(torch.equal(back, host))))
ctx.close()
