#!/usr/bin/env python
"""Run one BASELINE.json config through the in-memory hot path on ONE GPU and print sizes, stage times and a few
size-independent invariants.  For the multi-GPU runs use bench.py under torchrun.

    python tools/run_config.py --config cfg3_4.5Gbp_tobacco --low-memory
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polycracker_b200 import binding, synth  # noqa: E402
from polycracker_b200.pipeline import ChunkLayout, HostArena, run_hot_path  # noqa: E402


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


HBM_PEAK_GBS = _hbm_peak()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3_4.5Gbp_tobacco", choices=sorted(synth.CONFIGS))
    ap.add_argument("--genome-size", type=int, default=0)
    ap.add_argument("--low-memory", action="store_true")
    ap.add_argument("--fetch", action="store_true", help="also copy the CSR + TF-IDF to the host")
    ap.add_argument("--part-mbytes", type=int, default=0)
    ap.add_argument("--k-list", default="", help="comma separated k values (BASELINE config 5: 15,23,31); default = the config's k")
    ap.add_argument("--tune", default="", help="pc_tune settings, e.g. part_mode=1")
    a = ap.parse_args()
    import torch
    cfg = dict(synth.CONFIGS[a.config])
    gsize = a.genome_size or cfg["genome_size"]
    params = synth.make_params(cfg["seed"], gsize, n_subgenomes=cfg["n_subgenomes"])
    t0 = time.time()
    pinned = torch.empty(gsize, dtype=torch.uint8, pin_memory=True)
    seq, off, names = synth.generate(params, out=pinned.numpy(), threads=min(16, os.cpu_count() or 8))
    t_gen = time.time() - t0
    layout = ChunkLayout(names, off, cfg["chunk_size"])
    layout.rows()
    ctx = binding.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if a.part_mbytes:
        ctx.tune(part_mbytes=a.part_mbytes)
    if a.tune:
        ctx.tune(**{kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.tune.split(",")})
    dev = pinned[:len(seq)].to("cuda")
    arena = HostArena()
    for k in ([int(x) for x in a.k_list.split(",")] if a.k_list else [cfg["k"]]):
        run_k(a, cfg, k, ctx, dev, off, names, layout, arena, seq, t_gen, torch)


def run_k(a, cfg, k, ctx, dev, off, names, layout, arena, seq, t_gen, torch):
    cfg = dict(cfg, k=k)
    out = {"k": k}
    for it in range(2):                                    # first pass allocates, second is the steady state
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = run_hot_path(ctx, dev, off, names, chunk_size=cfg["chunk_size"], k=cfg["k"],
                           kmer_low_count=cfg["kmer_low_count"], fetch=a.fetch, layout=layout, arena=arena,
                           low_memory=a.low_memory)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out["pass%d_s" % it] = round(dt, 4)
    st = res.stats
    det, flags = st.get("count_detail", {}), st.get("count_flags", {})
    out["count"] = det
    # BASELINE configs[4]: hash-table occupancy and achieved HBM GB/s.  The count has no global table: every sub-bucket is
    # counted in its own shared-memory table, so "occupancy" = distinct k-mers / (tables x slots), plus the sub-buckets whose
    # k-mers did not fit and went through the global-table fall-back.
    if det.get("kind") == "list" and det.get("sub_buckets"):
        slots = flags.get("slots_per_table", 8192)
        out["occupancy"] = {"shared_tables": det["sub_buckets"], "slots_per_table": slots,
                            "mean_load": round(st["distinct"] / (det["sub_buckets"] * slots), 4),
                            "overflowed_sub_buckets": det["overflowed"],
                            "overflowed_fraction": round(det["overflowed"] / det["sub_buckets"], 5),
                            "largest_unit_kmers": flags.get("largest_unit"), "kept_kmers": det["kept"],
                            "super_kmer_records": det.get("skm_records"), "minimizer_length": det.get("skm_m")}
    tm = res.timings
    G = sum(int(l // 32 + 1) for l in (layout.chunk_end - layout.chunk_start).tolist()) * 32
    T, nnz, Ds, R = st["valid_windows"], st["nnz"], st["n_selected"], st["n_rows"]
    k3_ms = sum(tm.get(x, 0.0) for x in ("part_hist", "part_scatter", "part2", "table_init", "expand", "count", "count_fallback"))
    units = {"K1_pack": (1.375 * G, tm.get("pack", 0.0)), "K3_count": (0.375 * G + 24.0 * T, k3_ms),
             "K5_probe": (0.375 * G + 12.0 * T + 8.0 * nnz, tm.get("probe", 0.0)),
             "K6_rows_to_csr": (16.0 * nnz + 8.0 * R, tm.get("row_sort", 0.0) + tm.get("emit", 0.0)),
             "K7_tfidf": (12.0 * nnz + 4.0 * Ds, tm.get("tfidf", 0.0))}
    out["achieved_gbs"] = {n: round(b / (ms * 1e-3) / 1e9, 1) for n, (b, ms) in units.items() if ms > 0}
    out["achieved_frac_of_hbm_peak"] = {n: round(v / HBM_PEAK_GBS, 3) for n, v in out["achieved_gbs"].items()}
    free_b, total_b = torch.cuda.mem_get_info()
    out.update(config=a.config, genome_bases=int(len(seq)), gen_s=round(t_gen, 1), gbp_per_s=round(len(seq) / dt / 1e9, 2),
               stats={k_: int(v) for k_, v in st.items() if isinstance(v, (int, np.integer)) and not isinstance(v, bool)},
               stage_ms={k_: round(v, 2) for k_, v in res.timings.items() if v > 0},
               hbm_used_gb=round((total_b - free_b) / 1e9, 1), hbm_total_gb=round(total_b / 1e9, 1),
               torch_peak_gb=round(torch.cuda.max_memory_allocated() / 1e9, 1))
    if a.fetch:
        m = res.csr_matrix()
        lens = dict(zip(layout.chunk_names, (layout.chunk_end - layout.chunk_start).tolist()))
        rs = np.asarray(m.sum(axis=1)).ravel()
        out["row_sums_bounded"] = bool(all(rs[i] <= lens[s] - cfg["k"] + 1 for i, s in enumerate(res.scaffolds)))
        out["indices_sorted"] = bool(all(np.all(np.diff(res.indices[res.indptr[r]:res.indptr[r + 1]]) > 0)
                                         for r in range(0, m.shape[0], max(1, m.shape[0] // 200))))
        out["keys_sorted_unique"] = bool(np.all(np.diff(res.kmer_keys.astype(np.uint64)) > 0))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
