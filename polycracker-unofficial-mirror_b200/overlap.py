"""Steady-state end-to-end runner: host -> device copy of the NEXT genome and device -> host copy of the PREVIOUS
result overlap the compute of the current step (three CUDA streams, double-buffered staging on both sides).

One pass of the hot path on a 500 Mbp genome is ~20 ms of kernels, but PCIe moves 0.5 GB in and 1.2 GB out (~33 ms at
50 GB/s): a caller that processes a series of genomes (or shards, or thresholds) back to back is PCIe-bound unless the
copies of neighbouring steps overlap.  PCIe is full duplex, so in steady state a step costs
max(compute, H2D, D2H) instead of their sum.

    pipe = StepPipeline(ctx, device)
    pipe.prefetch(pinned_genome_0)
    for i in range(n):
        if i + 1 < n: pipe.prefetch(pinned_genome_{i+1})     # async, side stream
        done = pipe.step(run)                                 # run(dev_genome) -> HotPathResult(fetch=False); returns the
                                                              # PREVIOUS step's host result (None for the first call)
    last = pipe.flush()

Every step's input really crosses PCIe and every step's CSR + TF-IDF really lands in page-locked host memory; only the
order of the copies relative to the neighbours' kernels changes.  PyTorch supplies streams, events and pinned memory;
all compute goes through the C ABI as everywhere else.
"""
import numpy as np

from .pipeline import HostArena, HotPathResult


class StepPipeline:
    def __init__(self, ctx, device, depth=2, sm_copy_ctas=0):
        import torch
        self.torch, self.ctx, self.device, self.depth = torch, ctx, device, depth
        # the context computes on a stream of this pipeline, so the events below order against its kernels (a context
        # given torch's legacy default stream handle, 0, would fall back to its private stream)
        self.s_main = torch.cuda.Stream(device)
        ctx.set_stream(self.s_main.cuda_stream)
        self.s_h2d = torch.cuda.Stream(device)
        self.s_d2h = torch.cuda.Stream(device)
        # Measured on B200 (tools/overlap_probe.py): a bulk H2D on a copy engine overlaps a pass perfectly (20.6 ms for
        # both); a bulk D2H on a copy engine did not (35 ms) although no kernel got slower -- the engine serves its queue
        # in order, and every tiny cudaMemcpy read-back of the context waited behind hundreds of megabytes.  The context
        # now reads its counters back through a device-mapped mailbox (api.cu read_small), which removes that coupling.
        # Copies made by SMs instead (sm_copy_ctas > 0, pc_copy_async) reach the same 52 GB/s but displace CTAs of the
        # persistent kernels (count 4.1 -> 6.7 ms): kept as an option, not the default.
        self.piece = 64 << 20
        self.sm_copy_ctas = sm_copy_ctas
        self.dev_in = [None] * depth                         # device staging of the inputs
        self.in_ready = [None] * depth                       # event: H2D of that slot finished
        self.in_len = [0] * depth
        self.snap = [dict() for _ in range(depth)]           # device snapshots of the results
        self.snap_free = [None] * depth                      # event: D2H out of that snapshot finished
        self.arena = [HostArena() for _ in range(depth)]
        self.n_prefetched = 0
        self.n_steps = 0
        self.pending = None                                  # (slot, event, result skeleton, host views)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # ---- input side ------------------------------------------------------------------------------------------
    def prefetch(self, host_tensor):
        """Start the H2D copy of the next step's genome (pinned torch uint8 tensor) on the copy-in stream."""
        torch = self.torch
        slot = self.n_prefetched % self.depth
        n = host_tensor.numel()
        if self.dev_in[slot] is None or self.dev_in[slot].numel() < n:
            self.dev_in[slot] = torch.empty(int(n * 1.05) + 256, dtype=torch.uint8, device=self.device)
        # the slot was last read by step (n_prefetched - depth), which has returned (host-synchronous C calls)
        with torch.cuda.stream(self.s_h2d):
            self._copy_pieces(self.dev_in[slot][:n], host_tensor)
            ev = torch.cuda.Event()
            ev.record(self.s_h2d)
        self.in_ready[slot], self.in_len[slot] = ev, n
        self.n_prefetched += 1
        self.h2d_bytes += n

    def _copy_pieces(self, dst, src):
        """Bulk copy on the CURRENT torch stream: by SMs (pc_copy_async) when sm_copy_ctas > 0, else by the copy engine
        in pieces."""
        if self.sm_copy_ctas > 0:
            self.ctx.copy_async(dst, src, stream=self.torch.cuda.current_stream(self.device).cuda_stream,
                                n_ctas=self.sm_copy_ctas)
            return
        n, step = src.numel(), max(1, self.piece // src.element_size())
        for a in range(0, n, step):
            dst[a:a + step].copy_(src[a:a + step], non_blocking=True)

    def close(self):
        """Hand the context back to its private stream."""
        self.s_main.synchronize()
        self.ctx.set_stream(None)                             # None = the context's own stream (0 would mean cudaStreamLegacy)

    def _dev(self, slot, name, n, dtype):
        torch = self.torch
        t = self.snap[slot].get(name)
        if t is None or t.numel() < max(n, 1) or t.dtype != dtype:
            t = self.snap[slot][name] = torch.empty(int(max(n, 1) * 1.25) + 16, dtype=dtype, device=self.device)
        return t[:n]

    # ---- one step ----------------------------------------------------------------------------------------------
    def step(self, run, tfidf=True, df=None):
        """run(device_genome_tensor) must execute one pass with fetch=False on the current stream and return its
        HotPathResult.  Returns the host-side result of the PREVIOUS step (None on the first call)."""
        torch = self.torch
        assert self.n_prefetched > self.n_steps, "prefetch() the input of this step first"
        slot = self.n_steps % self.depth
        with torch.cuda.stream(self.s_main):
            return self._step(run, tfidf, df, slot)

    def _step(self, run, tfidf, df, slot):
        torch = self.torch
        main = self.s_main
        main.wait_event(self.in_ready[slot])
        res = run(self.dev_in[slot][:self.in_len[slot]])
        st = res.stats
        n_sel, nnz, n_rows = int(st["n_selected"]), int(st["nnz"]), int(st["n_rows"])
        # device snapshot on the compute stream (the context's own buffers are overwritten by the next step)
        if self.snap_free[slot] is not None:
            main.wait_event(self.snap_free[slot])
        d_keys = self._dev(slot, "keys", n_sel, torch.int64)
        d_indptr = self._dev(slot, "indptr", n_rows + 1, torch.int64)
        d_df = self._dev(slot, "df", n_sel, torch.int32)
        if n_sel:
            self.ctx.selected_into(d_keys)
            if df is None:
                self.ctx.df_into(d_df)
            else:
                d_df.copy_(df[:n_sel])
        # densest hand-off (pc_csr_get8 / pc_tfidf_factors_get): one byte per column gap, one per count, idf + row norms
        # instead of the tf-idf array -- 2 bytes per matrix entry cross PCIe instead of 12; the host rebuilds the int32 /
        # float32 arrays bit for bit (HotPathResult.indices / .data / .tfidf)
        gdt = self.ctx.gap_dtype(n_sel, n_rows, nnz)         # one-byte column gaps, two for a large column universe
        wide = gdt == np.uint16
        d_dcol = self._dev(slot, "dcol16" if wide else "dcol8", nnz, torch.int16 if wide else torch.uint8)
        d_cnt = None if wide else self._dev(slot, "cnt8", nnz, torch.uint8)   # wide: gap << 4 | count in one uint16
        n_col, n_cnt = self.ctx.csr8_into(d_indptr, d_dcol if (nnz or wide) else None, d_cnt if nnz else None)
        # the escape lists (a few MB) leave with the bulk copies on the copy-out stream: a host-synchronous read here would
        # queue on the copy engine behind the previous step's D2H and stall the pipeline
        d_ecp, d_ecv = self._dev(slot, "esc_col_pos", n_col, torch.int64), self._dev(slot, "esc_col_val", n_col, torch.int32)
        d_ekp, d_ekv = self._dev(slot, "esc_cnt_pos", n_cnt, torch.int64), self._dev(slot, "esc_cnt_val", n_cnt, torch.float32)
        if n_col or n_cnt:
            self.ctx.csr8_escapes_into(d_ecp if n_col else None, d_ecv if n_col else None, d_ekp if n_cnt else None,
                                       d_ekv if n_cnt else None)
        d_idf = d_norm = None
        if tfidf:
            d_idf = self._dev(slot, "idf", n_sel, torch.float32)
            d_norm = self._dev(slot, "row_norm", n_rows, torch.float64)
            self.ctx.tfidf_factors_into(d_idf if n_sel else None, d_norm if n_rows else None)
        ev = torch.cuda.Event()
        ev.record(main)
        # D2H on the copy-out stream into this slot's page-locked arena
        a = self.arena[slot]
        h = dict(keys=a.get("keys", n_sel, np.uint64), indptr=a.get("indptr", n_rows + 1, np.int64),
                 dcol8=a.get("dcol16" if wide else "dcol8", nnz, gdt), cnt8=None if wide else a.get("cnt8", nnz, np.uint8),
                 idf=a.get("idf", n_sel, np.float32) if tfidf else None,
                 row_norm=a.get("row_norm", n_rows, np.float64) if tfidf else None, df=a.get("df", n_sel, np.int32),
                 esc_col_pos=a.get("esc_col_pos", n_col, np.int64), esc_col_val=a.get("esc_col_val", n_col, np.int32),
                 esc_cnt_pos=a.get("esc_cnt_pos", n_cnt, np.int64), esc_cnt_val=a.get("esc_cnt_val", n_cnt, np.float32))
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(ev)
            for name, src in (("keys", d_keys), ("indptr", d_indptr), ("dcol8", d_dcol), ("cnt8", d_cnt),
                              ("idf", d_idf), ("row_norm", d_norm), ("df", d_df), ("esc_col_pos", d_ecp),
                              ("esc_col_val", d_ecv), ("esc_cnt_pos", d_ekp), ("esc_cnt_val", d_ekv)):
                if src is None or src.numel() == 0:
                    continue
                hv = h[name]
                dst = torch.from_numpy(hv.view(np.int64) if name == "keys" else hv.view(np.int16) if hv.dtype == np.uint16 else hv)
                self._copy_pieces(dst, src)
                self.d2h_bytes += src.numel() * src.element_size()
            done = torch.cuda.Event()
            done.record(self.s_d2h)
        self.snap_free[slot] = done
        previous = self._collect()
        self.pending = (done, res, h)
        self.n_steps += 1
        return previous

    def _collect(self):
        """Host result of the pending step.  Its arrays are views into this pipeline's page-locked arena: they stay valid
        until the step() after next reuses the slot (depth = 2) -- copy what must live longer."""
        if self.pending is None:
            return None
        done, res, h = self.pending
        done.synchronize()
        self.pending = None
        escapes8 = ((h["esc_col_pos"], h["esc_col_val"]), (h["esc_cnt_pos"], h["esc_cnt_val"]))
        return HotPathResult(res.layout, res.k, res.scaffolds, h["keys"], h["indptr"], None, None, None, h["df"],
                             res.stats, res.timings, idf=h["idf"], row_norm=h["row_norm"], dcol8=h["dcol8"], cnt8=h["cnt8"],
                             escapes8=escapes8)

    def flush(self):
        """Host result of the last step."""
        return self._collect()
