#!/usr/bin/env python
"""N3 Gram matrix at the bench workload: the hand-written tcgen05 kernel against the library GEMM (float32 and TF32).
    python tools/gram_bench.py [genome_size]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from polycracker_b200 import binding, synth, gram
from polycracker_b200.pipeline import ChunkLayout, run_hot_path

size = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
cfg = synth.CONFIGS["cfg2_500Mbp_allotetraploid"]
p = synth.make_params(cfg["seed"], size, n_subgenomes=cfg["n_subgenomes"])
seq, off, names = synth.generate(p, threads=8)
lay = ChunkLayout(names, off, 75000)
ctx = binding.Context(0)
res = run_hot_path(ctx, torch.from_numpy(seq).cuda(), off, names, chunk_size=75000, k=23, kmer_low_count=30, fetch=False, layout=lay)
R, D, nnz = res.stats["n_rows"], res.stats["n_selected"], res.stats["nnz"]
out = {"rows": R, "cols": D, "nnz": nnz, "flop_full": 2.0 * R * R * D}
ref = None
for name, kw in (("tcgen05_tf32", dict(engine="tcgen05")), ("cublas_tf32", dict(engine="cublas", tf32=True)),
                 ("cublas_fp32", dict(engine="cublas"))):
    ts = []
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        K = gram.gram_matrix(ctx, R, D, values="tfidf", kernel="linear", **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = min(ts)
    out[name] = {"s": round(t, 4), "tflops_full_matrix": round(2.0 * R * R * D / t / 1e12, 1)}
    if name == "cublas_fp32":
        ref = K
    else:
        out[name]["K"] = K
for name in ("tcgen05_tf32", "cublas_tf32"):
    K = out[name].pop("K")
    out[name]["max_rel_err_vs_fp32"] = float(((K - ref).abs() / ref.abs().clamp_min(1e-6)).max())
print(json.dumps(out))
