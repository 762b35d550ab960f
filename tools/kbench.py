#!/usr/bin/env python
"""Kernel tuning harness (development tool): times pc_count / pc_build_matrix variants on one synthetic genome.
    python tools/kbench.py --genome-size 200000000 --what count
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding, synth  # noqa: E402
from polycracker_b200.pipeline import ChunkLayout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-size", type=int, default=200_000_000)
    ap.add_argument("--k", type=int, default=23)
    ap.add_argument("--what", default="count")
    ap.add_argument("--variants", default="")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--low", type=int, default=30, help="repeat threshold for --what matrix")
    a = ap.parse_args()
    import torch
    p = synth.make_params(20260002, a.genome_size)
    seq, off, names = synth.generate(p, threads=8)
    lay = ChunkLayout(names, off, 75000)
    scaffolds, chunk_row = lay.rows()
    dev = torch.from_numpy(seq).cuda()
    ctx = binding.Context(0)
    ctx.load_sequences(dev, lay.chunk_begin_abs, lay.chunk_end_abs)
    base = None
    if a.what == "count":
        variants = []
        if a.variants:
            for v in a.variants.split(";"):
                variants.append(dict(kv.split("=") for kv in v.split(",")))
        else:
            for mode, batch, minb in itertools.product((0, 1, 2), (8, 16, 4), (2, 3, 4)):
                if (batch == 16 and minb != 2) or (batch == 4 and minb != 4):
                    continue
                variants.append(dict(count_mode=mode, count_batch=batch, count_minb=minb))
            variants += [dict(count_mode=1, count_batch=8, count_minb=2, l2_fetch_granularity=g) for g in (32, 128, 64)]
            variants += [dict(count_mode=1, count_batch=8, count_minb=2, table_load_pct=l) for l in (40, 50, 70, 80, 60)]
        for v in variants:
            ctx.tune(**{k_: int(x) for k_, x in v.items()})
            ts = []
            for _ in range(a.reps):
                ctx.reset_counters()
                ctx.count(a.k)
                t = ctx.timings()
                ctx.select(max(a.low, int(v.get("keep_min", 1))))
                t2 = ctx.timings()
                ts.append((t["count"] + t["part_hist"] + t["part_scatter"] + t["part2"] + t["table_init"] + t["count_fallback"] + t["expand"] + t2["select"],
                           t["table_init"], t["count"], t["part_hist"], t["part_scatter"], t["part2"], t2["select"], t["count_fallback"], t["expand"]))
            info = ctx.count_info()
            det = ctx.count_detail()
            sig = (info["distinct"], info["valid_windows"])
            if base is None:
                base = sig
            best = min(ts)
            print(json.dumps(dict(variant=v, total_ms=round(best[0], 3), init_ms=round(best[1], 3), count_ms=round(best[2], 3),
                                  hist_ms=round(best[3], 3), scatter_ms=round(best[4], 3), part2_ms=round(best[5], 3),
                                  select_ms=round(best[6], 3), fallback_ms=round(best[7], 3), expand_ms=round(best[8], 3), detail=det,
                                  ginserts_per_s=round(info["valid_windows"] / best[0] / 1e6, 2),
                                  load=round(info["distinct"] / info["capacity"], 3), max_probe=info["max_probe"],
                                  ok=(sig == base))), flush=True)
    elif a.what == "matrix":
        ctx.count(a.k)
        variants = [dict()]
        if a.variants:
            variants = [dict(kv.split("=") for kv in v.split(",")) for v in a.variants.split(";")]
        for v in variants:
            if v:
                ctx.tune(**{k_: int(x) for k_, x in v.items()})
                if "hash_mode" in v:
                    ctx.count(a.k)
            best = None
            for _ in range(a.reps):
                n_sel = ctx.select(a.low)
                nnz = ctx.build_matrix(chunk_row, len(scaffolds))
                ctx.tfidf(len(scaffolds))
                t = ctx.timings()
                cur = {k_: round(t[k_], 3) for k_ in ("select", "sort", "rep_table", "probe", "row_sort", "emit", "tfidf")}
                if best is None or cur["probe"] < best["probe"]:
                    best = cur
            indptr, indices, data = ctx.csr(len(scaffolds), nnz)
            sig = (n_sel, nnz, int(indices.astype(np.int64).sum()), float(data.sum()))
            if base is None:
                base = sig
            print(json.dumps(dict(variant=v, n_sel=n_sel, nnz=nnz, ok=(sig == base), **best)), flush=True)


if __name__ == "__main__":
    main()
