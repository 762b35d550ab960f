#!/usr/bin/env python
"""Stage timeline of one resident step of the bench workload: where the device idles between stages."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import ChunkLayout, run_hot_path

size = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
cfg = synth.CONFIGS["cfg2_500Mbp_allotetraploid"]
p = synth.make_params(cfg["seed"], size, n_subgenomes=cfg["n_subgenomes"])
seq, off, names = synth.generate(p, threads=8)
lay = ChunkLayout(names, off, 75000); lay.rows()
dev = torch.from_numpy(seq).cuda()
ctx = binding.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
steps = []
for i in range(int(os.environ.get('PC_TL_ITERS', '4'))):
    ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    res = run_hot_path(ctx, dev, off, names, chunk_size=75000, k=23, kmer_low_count=30, fetch=False, layout=lay)
    e1.record(); torch.cuda.synchronize()
    tl = ctx.stage_timeline()
    steps.append(round(e0.elapsed_time(e1), 2))
print("step ms", e0.elapsed_time(e1), steps)
prev = None; busy = 0.0
for name, a, b in tl:
    gap = (a - prev) if prev is not None else 0.0
    print("%-16s start %8.3f  end %8.3f  dur %7.3f  gap before %7.3f" % (name, a, b, b - a, gap))
    prev = b; busy += b - a
print("sum of stages %.3f, span %.3f" % (busy, tl[-1][2] - tl[0][1]))
