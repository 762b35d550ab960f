import sys, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import ChunkLayout
g = int(sys.argv[1])
p = synth.make_params(20260003, g); 
pinned = torch.empty(g, dtype=torch.uint8, pin_memory=True)
seq, off, names = synth.generate(p, out=pinned.numpy(), threads=16)
lay = ChunkLayout(names, off, 75000)
def used(tag):
    f, t = torch.cuda.mem_get_info(); print(tag, round((t - f) / 1e9, 2), "GB used of", round(t / 1e9, 1), flush=True)
used("start")
ctx = binding.Context(0); used("ctx")
dev = pinned[:len(seq)].to("cuda"); used("dev genome")
ctx.load_sequences(dev, lay.chunk_begin_abs, lay.chunk_end_abs); used("loaded")
try:
    ctx.count(23)
except Exception as e:
    print("ERR", e)
used("counted")
print(ctx.count_info())
