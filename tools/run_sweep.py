#!/usr/bin/env python
"""BASELINE.json config 4 (and 3): a FIXED-size genome sharded over all ranks (strong scaling), one count, then a sweep
of repeat-count thresholds that re-uses the owners' count tables (only K4 -> all-gather -> K5..K7 are repeated).

    torchrun --nproc-per-node 8 tools/run_sweep.py --config cfg4_15Gbp_wheat --thresholds 10,30,100,300,1000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4_15Gbp_wheat")
    ap.add_argument("--genome-size", type=int, default=0)
    ap.add_argument("--thresholds", default="10,30,100,300,1000")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from polycracker_b200 import binding, distributed as pcd, synth
    from polycracker_b200.pipeline import ChunkLayout
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    cfg = dict(synth.CONFIGS[a.config])
    params = synth.make_params(cfg["seed"], a.genome_size or cfg["genome_size"], n_subgenomes=cfg["n_subgenomes"])
    lay_all = synth.layout(params)
    names = [n for n, _ in lay_all]
    offs = np.concatenate([[0], np.cumsum([l for _, l in lay_all])]).astype(np.int64)
    layout = ChunkLayout(names, offs, cfg["chunk_size"])
    blocks = pcd.shard_chunks(layout, world)
    ids, begins, ends, local_bases = pcd.local_buffer_plan(layout, blocks[rank])
    t0 = time.time()
    pinned = torch.empty(max(local_bases, 1), dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    for c in sorted(set(layout.chunk_seq[ids].tolist())):
        chrom, _, _ = synth.generate(params, chroms=[c], threads=4)
        for j in np.flatnonzero(layout.chunk_seq[ids] == c).tolist():
            i = ids[j]
            host[begins[j]:ends[j]] = chrom[layout.chunk_start[i]:layout.chunk_end[i]]
        del chrom
    t_gen = time.time() - t0
    dev = pinned[:local_bases].to(device)
    ctx = binding.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    layout.rows()
    thresholds = [int(x) for x in a.thresholds.split(",")]
    results = []
    for it in range(2):                                    # pass 0 allocates, pass 1 is reported
        results = []
        for ti, low in enumerate(thresholds):
            dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
            if ti == 0:
                res = pcd.run_hot_path_distributed(ctx, layout, None, device, chunk_size=cfg["chunk_size"], k=cfg["k"],
                                                   kmer_low_count=low, fetch=False, staged=(dev, ids, begins, ends))
            else:
                res = pcd.rerun_threshold(ctx, layout, ids, device, k=cfg["k"], kmer_low_count=low, fetch=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if ti == 0:
                first = res
            results.append(dict(threshold=low, seconds=round(float(t.item()), 4), n_selected=res.stats["n_selected"],
                                nnz_global=res.stats["nnz_global"], rows=res.stats["n_rows_global"],
                                full_path=(ti == 0), phase_ms=res.stats.get("phase_ms")))
    if rank == 0:
        free_b, total_b = torch.cuda.mem_get_info()
        print(json.dumps(dict(config=a.config, n_gpus=world, genome_bases=int(offs[-1]), gen_s=round(t_gen, 1),
                              valid_windows=first.stats["valid_windows_global"], distinct=first.stats["distinct_global"],
                              gbp_per_s_full_path=round(int(offs[-1]) / results[0]["seconds"] / 1e9, 2),
                              hbm_used_gb_rank0=round((total_b - free_b) / 1e9, 1), sweep=results)))
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
