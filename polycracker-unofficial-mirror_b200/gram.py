"""Scope row N3, second half: the Gram matrix behind KernelPCA(kernel='linear' | 'cosine') (polycracker.py:387, 398-399).

sklearn forms K = X X^T (linear_kernel) or K = Xn Xn^T with L2-normalised rows (cosine_similarity) from the sparse
R x D matrix and hands it to an eigen-solver; this module forms the same K on the GPU from the matrix the context just
built, without the matrix ever leaving the device:

    for every panel of `panel_cols` columns:   P = dense float32 panel (pc_densify_panel, hand-written kernel)
                                               K += P P^T                   (pc_gram_accumulate: csrc/gram_tc.cuh)

The contraction is a hand-written sm_100a kernel (engine="tcgen05", the default): TMA-fed tcgen05.mma kind::tf32 with the
accumulator in TMEM, tiles on and above the diagonal only, K within 1e-3 relative of sklearn.  engine="cublas" keeps the
plain library GEMM (float32, or TF32 with tf32=True) as the cross-check.  X has ~2 % of its entries stored, so panels are
built on the fly (108 MB each at the bench workload: L2 resident while the tiles re-read them) instead of materialising
8.9 GB of dense X.
"""
import numpy as np


def gram_matrix(ctx, n_rows, n_cols, values="tfidf", kernel="linear", panel_cols=4096, tf32=False, device=None,
                engine="tcgen05"):
    """-> torch float32 CUDA tensor [n_rows, n_rows].  `ctx` holds a built matrix (and TF-IDF / scaled values when asked
    for).  kernel='cosine' divides by the row norms (rows of norm 0 stay 0), like sklearn's cosine_similarity."""
    import torch
    device = device or torch.device("cuda", ctx.device)
    stream = torch.cuda.Stream(device)
    ctx_stream_before = None
    K = torch.zeros((n_rows, n_rows), dtype=torch.float32, device=device)
    if n_rows == 0 or n_cols == 0:
        return K
    panel_cols = int(min(panel_cols, n_cols))
    if engine == "tcgen05":
        panel_cols = (panel_cols + 3) & ~3                     # TMA: row pitch a multiple of 16 bytes
    elif engine != "cublas":
        raise ValueError("engine must be 'tcgen05' or 'cublas'")
    bufs = [torch.empty((n_rows, panel_cols), dtype=torch.float32, device=device) for _ in range(2)]
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    if engine == "tcgen05":
        ctx.tune(gram_preround=1)                              # the panel kernel rounds to TF32: no pass of its own
    try:
        with torch.cuda.stream(stream):
            ctx.set_stream(stream.cuda_stream)               # panel kernels and GEMMs are ordered on one stream
            for i, col0 in enumerate(range(0, n_cols, panel_cols)):
                nc = min(panel_cols, n_cols - col0)
                P = bufs[i & 1]
                ctx.densify_panel(values, col0, nc, P)
                if engine == "tcgen05":
                    ctx.gram_accumulate(P, nc, K)
                else:
                    K.addmm_(P[:, :nc], P[:, :nc].t())
            if engine == "tcgen05":
                ctx.gram_finish(K)
            if kernel == "cosine":
                d = torch.sqrt(torch.clamp(torch.diagonal(K), min=0))
                inv = torch.where(d > 0, 1.0 / d, torch.zeros_like(d))
                K *= inv[:, None]
                K *= inv[None, :]
            elif kernel != "linear":
                raise ValueError("kernel must be 'linear' or 'cosine'")
        stream.synchronize()
    finally:
        if engine == "tcgen05":
            ctx.tune(gram_preround=0)
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        ctx.set_stream(ctx_stream_before)
    return K


def gram_matrix_numpy(*a, **kw):
    return gram_matrix(*a, **kw).cpu().numpy()
