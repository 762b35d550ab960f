#!/usr/bin/env python
"""One or more passes of the hot path on a synthetic cfg2-shaped genome (profiling target for ncu).
    python tools/prof_step.py --genome-size 200000000 --passes 1 [--tune part_mode=1,...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding, synth  # noqa: E402
from polycracker_b200.pipeline import ChunkLayout, run_hot_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-size", type=int, default=200_000_000)
    ap.add_argument("--k", type=int, default=23)
    ap.add_argument("--low", type=int, default=30)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--tune", default="")
    a = ap.parse_args()
    import torch
    p = synth.make_params(20260002, a.genome_size)
    seq, off, names = synth.generate(p, threads=8)
    lay = ChunkLayout(names, off, 75000)
    dev = torch.from_numpy(seq).cuda()
    ctx = binding.Context(0)
    if a.tune:
        ctx.tune(**{kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.tune.split(",")})
    for _ in range(a.passes):
        res = run_hot_path(ctx, dev, off, names, chunk_size=75000, k=a.k, kmer_low_count=a.low, fetch=False, layout=lay)
    print(json.dumps(dict(stats=res.stats, timings={k_: round(v, 3) for k_, v in res.timings.items() if v}, detail=ctx.count_detail())))


if __name__ == "__main__":
    main()
