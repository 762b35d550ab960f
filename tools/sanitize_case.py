"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): direct and partitioned count, stream and roller
probe, K6, K7, on a 600 kb genome with N runs and odd chunk sizes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import run_hot_path

p = synth.small_params(seed=5, genome_size=600_000, n_fraction=0.01)
seq, off, names = synth.generate(p)
with binding.Context(0) as ctx:
    for mode, chunk in ((0, 9973), (1, 20000)):
        ctx.tune(part_mode=mode, part_mbytes=1)
        r = run_hot_path(ctx, seq, off, names, chunk_size=chunk, k=23, kmer_low_count=5)
        print("mode", mode, "shape", r.shape, "nnz", len(r.data), "distinct", r.stats["distinct"])
    ctx.tune(part_mode=1, probe_mode=0)
    r = run_hot_path(ctx, seq, off, names, chunk_size=20000, k=31, kmer_low_count=4)
    print("roller probe", r.shape, len(r.data))
print("done")
