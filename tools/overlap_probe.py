#!/usr/bin/env python
"""Does a bulk PCIe copy on a side stream overlap one pass of the hot path? (development tool)"""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding, synth  # noqa: E402
from polycracker_b200.pipeline import ChunkLayout, run_hot_path  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
p = synth.make_params(20260002, size)
seq, off, names = synth.generate(p, threads=8)
lay = ChunkLayout(names, off, 75000)
lay.rows()
dev = torch.from_numpy(seq).cuda()
ctx = binding.Context(0)
s_main, s_copy = torch.cuda.Stream(), torch.cuda.Stream()
ctx.set_stream(s_main.cuda_stream)
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
big = torch.empty(n, dtype=torch.uint8, device="cuda")


def compute():
    with torch.cuda.stream(s_main):
        run_hot_path(ctx, dev, off, names, chunk_size=75000, k=23, kmer_low_count=30, fetch=False, layout=lay)


def copy(kind, direction):
    dst, src = (host, big) if direction == "d2h" else (big, host)
    with torch.cuda.stream(s_copy):
        if kind == "ce":
            dst.copy_(src, non_blocking=True)
        elif kind == "ce_pieces":
            for a in range(0, n, 4 << 20):
                dst[a:a + (4 << 20)].copy_(src[a:a + (4 << 20)], non_blocking=True)
        else:
            ctx.copy_async(dst, src, stream=s_copy.cuda_stream, n_ctas=int(kind[2:]))


def wall(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3


for _ in range(3):
    compute()
print("compute alone      %.1f ms" % min(wall(compute) for _ in range(3)))
def stages():
    t = ctx.timings()
    return sum(t.values()), {k_: round(v, 2) for k_, v in t.items() if v > 0.3}


print("stage sum alone", stages())
for direction in ("d2h", "h2d"):
    for kind in ("ce", "sm2", "sm4", "sm8"):
        alone = min(wall(lambda: copy(kind, direction)) for _ in range(2))
        both = min(wall(lambda: (copy(kind, direction), compute())) for _ in range(2))
        print("%s %-9s copy alone %.1f ms (%.1f GB/s)   copy + compute %.1f ms   kernel stage sum %.1f ms %s" % (
            direction, kind, alone, n / alone / 1e6, both, stages()[0], stages()[1] if kind in ("ce", "sm8") else ""))
ctx.close()
