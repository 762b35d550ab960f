"""Multi-GPU hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing.

Sharding: chunks are independent records, so the string-sorted chunk list is cut into `world`
contiguous blocks balanced by bases; shard-local CSR blocks concatenate into the final row order.  The only global
coupling is the k-mer count reduction and the broadcast of the selected repeat set:

  1. local K1 + pc_partition: canonical k-mers of the local chunks, radix-partitioned by the top bits of their hash;
     rank r owns a contiguous range of buckets, so what each peer must receive is one contiguous slice
  2. all-to-all of the raw 8-byte k-mers (bucket sizes first) -> every k-mer instance is counted exactly once, by its
     owner, with the L2-resident bucket kernel (pc_count_buckets); owners threshold (K4) their share
  3. all-gather of the selected keys        -> every rank installs the identical sorted repeat set
  4. local probe + CSR (K5, K6); df all-reduce; TF-IDF (K7) with the global idf
  5. optional gather of the row blocks to rank 0 (host side; outside the timed region of bench.py)

The reference's only precedent is query-sharded bbmap with file concatenation (polycracker.nf:247-295).
"""
import numpy as np

from .pipeline import ChunkLayout, HotPathResult, dump_floor, effective_low_count, fetch_result


def bind_to_gpu_numa(gpu_index):
    """Pin this process (and therefore the pages of the pinned host buffers it allocates afterwards: first touch) to the
    CPUs that are local to its GPU.  With one rank per GPU on a two-socket host, unbound ranks put most staging buffers on
    one socket and every copy of the other socket's GPUs crosses the inter-socket link.  Returns a description or None
    when NVML / sched_setaffinity are unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or None
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d cpus local to gpu %d (%d-%d)" % (len(cpus), gpu_index, min(cpus), max(cpus))
    except Exception:
        return None


def _to_host(ctx, t, torch, device):
    """Small tensor -> numpy.  On CUDA through the context's mailbox (Context.to_host): a plain .cpu() is a copy-engine
    transfer and would wait behind the bulk D2H of a neighbouring pipeline step (overlap.py).  The context stream is
    torch's current stream here, so the read is ordered after the collective that produced `t`."""
    if ctx is not None and hasattr(ctx, "to_host") and getattr(device, "type", "cpu") == "cuda" and t.is_cuda:
        return ctx.to_host(t)
    return t.cpu().numpy()


def _select_raw(ctx, low, high, use_high):
    """The owner's share of the repeat set, unsorted and without a repeat table of its own: every rank sorts the all-gathered
    set in pc_set_selected anyway."""
    raw = hasattr(ctx, "tune")
    if raw:
        try:
            ctx.tune(select_raw=1)
        except Exception:
            raw = False                                        # (stand-in contexts of the CPU tests)
    try:
        return ctx.select(low, high, use_high, 1)
    finally:
        if raw:
            ctx.tune(select_raw=0)


def shard_chunks(layout, world):
    """Contiguous blocks of the string-sorted chunk list, balanced by bases.  Returns a list (per rank) of chunk
    index arrays, each in sorted-name order."""
    order = sorted(range(layout.n_chunks), key=lambda i: layout.chunk_names[i])
    lens = (layout.chunk_end - layout.chunk_start)[order] if order else np.zeros(0, np.int64)
    cum = np.cumsum(lens)
    total = int(cum[-1]) if len(cum) else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(order))
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    order = np.asarray(order, dtype=np.int64)
    return [order[cuts[r]:cuts[r + 1]] for r in range(world)]


def local_buffer_plan(layout, my_chunks):
    """Byte plan of this rank's private ASCII buffer: its chunks in (sequence, start) order, back to back.
    Returns (chunk ids in buffer order, begin offsets, end offsets, total bytes)."""
    ids = np.sort(np.asarray(my_chunks, dtype=np.int64))      # chunk table order == (sequence, start) order
    lens = layout.chunk_end[ids] - layout.chunk_start[ids]
    ends = np.cumsum(lens)
    begins = ends - lens
    return ids, begins, ends, int(ends[-1]) if len(ends) else 0


def _all_to_all_v(dist, send, send_counts, torch):
    """Variable-size all-to-all of a 1-D tensor: exchange sizes first, then the payload (one
    ncclGroupStart{Send/Recv} under NCCL)."""
    world = dist.get_world_size()
    sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc)
    recv_counts = rc.tolist()
    recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts))
    return recv, recv_counts


def _all_gather_v(dist, local, torch, ctx=None):
    """Variable-size all-gather of a 1-D tensor: counts first, then padded payloads."""
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    ns = [torch.empty(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(ns, n)
    counts = [int(x) for x in _to_host(ctx, torch.cat(ns), torch, local.device).tolist()]
    m = max(counts + [1])
    pad = torch.zeros(m, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    outs = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:c] for o, c in zip(outs, counts)]), counts


def plan_partition_bits(windows_global, world, part_bytes=64 << 20, table_load=0.6, slot_bytes=12, smem_mean=0):
    """Bucket bits shared by all ranks, at least one bucket per rank.  Table count: enough buckets that one sub-table
    (the unit that must stay L2 resident while it is being counted) is about part_bytes.  Shared-memory count
    (smem_mean = mean k-mers per sub-bucket > 0): the two-level split needs windows / smem_mean sub-buckets in total and
    level 1 takes the smaller half of the bits, as pc_count does on one GPU."""
    pb = 1
    if smem_mean > 0:
        nb_total = max(1, windows_global // smem_mean)
        while pb < 12 and (2 << (2 * pb)) < nb_total:
            pb += 1
    else:
        cap_bytes = windows_global / table_load * slot_bytes
        while pb < 12 and cap_bytes / (1 << pb) > part_bytes:
            pb += 1
    while (1 << pb) < world:
        pb += 1
    return pb


def owner_ranges(P, world):
    """rank r owns the level-1 buckets [q_lo[r], q_lo[r + 1])"""
    return [(r * P + world - 1) // world for r in range(world + 1)]


def owner_segments(sizes_all, q_lo, rank):
    """sizes_all[sender, bucket] = units (k-mers or records) every sender holds per bucket, bucket-sorted in its buffer.
    -> (seg_begin, seg_len), both [my bucket, sender]: where the owner `rank` finds each sender's slice of its buckets
    inside that sender's buffer."""
    sizes_all = np.asarray(sizes_all, dtype=np.int64)
    sender_off = np.cumsum(sizes_all, axis=1) - sizes_all
    lo, hi = q_lo[rank], q_lo[rank + 1]
    return sender_off[:, lo:hi].T.copy(), sizes_all[:, lo:hi].T.copy()


def bind_context_stream(ctx, torch, device):
    """The context must compute on torch's current stream: its small read-backs (Context.to_host) are ordered on the context
    stream, so only then do they wait for the collective that produced their input.  Idempotent."""
    if getattr(device, "type", "cpu") == "cuda" and hasattr(ctx, "set_stream"):
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def device_view(torch, ptr, n, device):
    """Zero-copy int64 torch view of n k-mers at a raw device pointer owned by the context."""
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DevView(ptr, n), device=device)


def peer_pointers(ctx, torch, dist, device, handle, own_ptr, cache_name="_pc_peers", handles=None):
    """Device pointers of every rank's exported buffer in THIS process (own buffer: own pointer).  The 64-byte CUDA IPC
    handles are all-gathered every step (a 64-byte collective -- or they arrive in `handles`, piggy-backed on a collective
    the caller needed anyway) and a peer buffer is (re)opened only when its handle changed, i.e. when that rank had to grow
    its buffer; the mapping it replaces is closed."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if handles is None:
        mine = torch.from_numpy(np.ascontiguousarray(handle).copy()).to(device)
        allh = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(allh, mine)
        handles = [h.tobytes() for h in _to_host(ctx, torch.stack(allh), torch, device)]
    cache = getattr(ctx, cache_name, None)
    if cache is None:
        cache = {}
        setattr(ctx, cache_name, cache)
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(int(own_ptr))
            continue
        ent = cache.get(r)
        if ent is None or ent[0] != handles[r]:
            if ent is not None and hasattr(ctx, "peer_close"):
                ctx.peer_close(ent[1])
            ent = cache[r] = (handles[r], ctx.peer_open(np.frombuffer(handles[r], dtype=np.uint8)))
        ptrs.append(ent[1])
    return ptrs


def _stream_sync(torch, device):
    """The C ABI is host-synchronous on the context's stream; collectives are asynchronous on torch's stream.  When
    the two streams differ, wait for the collective before handing its output to the context (and vice versa the
    context calls have already synchronised).  A no-op on CPU (gloo)."""
    if getattr(device, "type", "cpu") == "cuda":
        torch.cuda.current_stream(device).synchronize()


def run_hot_path_distributed(ctx, layout, fetch_chunks, device, chunk_size=75000, k=23, kmer_low_count=100,
                             high_bool=0, kmer_high_count=2000000, sampling_sensitivity=1, remove_non_chunk=1,
                             min_chunk_threshold=0, min_chunk_size=0, prefilter=0, tfidf=True, fetch=True,
                             staged=None, arena=None, exchange="auto"):
    """One rank's share of the hot path.  `layout` describes the WHOLE genome (every rank builds the same one from
    names + lengths); `fetch_chunks(ids, begins, ends, total)` returns this rank's ASCII buffer (uint8, host numpy /
    pinned torch / CUDA torch) laid out by local_buffer_plan; `staged` short-circuits it with a prebuilt
    (buffer, ids, begins, ends).  Returns a HotPathResult for the local row block (stats carry the global sizes).
    """
    import time
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    bind_context_stream(ctx, torch, device)
    phase, t_last = {}, [time.perf_counter()]

    def mark(name):
        """host wall time of a phase, device work included (every context call is host-synchronous; collectives are
        followed by a stream sync)"""
        _stream_sync(torch, device)
        now = time.perf_counter()
        phase[name] = phase.get(name, 0.0) + (now - t_last[0]) * 1e3
        t_last[0] = now

    if staged is None:
        blocks = shard_chunks(layout, world)
        ids, begins, ends, total = local_buffer_plan(layout, blocks[rank])
        buf = fetch_chunks(ids, begins, ends, total)
    else:
        buf, ids, begins, ends = staged
    # global row universe (host integer logic, identical on every rank) and this rank's slice of it
    scaffolds_all, chunk_row_all = layout.rows(remove_non_chunk, min_chunk_threshold, min_chunk_size, prefilter)
    my_rows_global = chunk_row_all[ids]
    kept = np.flatnonzero(my_rows_global >= 0)
    row0 = int(my_rows_global[kept].min()) if len(kept) else 0
    local_chunk_row = np.where(my_rows_global >= 0, my_rows_global - row0, -1).astype(np.int32)
    n_local_rows = len(kept)
    assert n_local_rows == 0 or int(my_rows_global[kept].max()) - row0 + 1 == n_local_rows, "row block not contiguous"

    # 1. local partition: bucket = top pb bits of the k-mer hash, the same pb on every rank
    windows_global = int(np.maximum(0, (layout.chunk_end - layout.chunk_start) - k + 1).sum())
    smem_mean = 0
    if hasattr(ctx, "sub_plan"):
        per = ctx.sub_plan(1 << 40, 1 << 20)                 # sub-buckets per bucket for 2^20 k-mers per bucket ...
        smem_mean = (1 << 20) // per if per > 0 else 0        # ... = the mean sub-bucket size (0: table count in use)
    # super-k-mer path (pc_tune part_mode 3 / auto for k >= 17): 16-byte records -- runs of windows that share a minimizer --
    # travel instead of one 8-byte k-mer per window, so the owners read ~4x fewer bytes from their peers
    skm = None
    if hasattr(ctx, "skm_plan") and exchange != "kmers":
        plan = ctx.skm_plan(k, windows_global, world)
        if int(plan[5]) == 1:
            skm = plan
    pb = int(skm[0]) if skm is not None else plan_partition_bits(windows_global, world, smem_mean=smem_mean)
    P = 1 << pb
    q_lo = owner_ranges(P, world)
    mark("host_plan")
    ctx.load_sequences(buf, begins, ends)
    on_gpu = getattr(device, "type", "cpu") == "cuda"
    if hasattr(ctx, "watch_threshold"):
        ctx.watch_threshold(0)
    if hasattr(ctx, "tune"):
        ctx.tune(keep_min=dump_floor(k))                                  # the owners keep what the reference's counter dumps
    if skm is not None:
        nb2 = int(skm[1])
        use_peer = exchange in ("peer", "auto") and on_gpu and hasattr(ctx, "skm_export")
        handle = ctx.skm_export(k, skm) if use_peer else None
        off, recs, hist = ctx.skm_partition(k, skm)
        mark("pack+partition")
        n_local_records = int(off[-1])
        # one all-gather carries the bucket sizes AND the 64-byte IPC handle of the record buffer (8 more int64), and the
        # histogram all-to-all below is queued behind it before the host looks at either: one read-back instead of three
        payload = np.diff(off).astype(np.int64)
        if use_peer:
            payload = np.concatenate([payload, np.frombuffer(np.ascontiguousarray(handle).tobytes(), dtype=np.int64)])
        sizes = torch.from_numpy(payload).to(device)
        gathered = [torch.empty(len(payload), dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(gathered, sizes)
        # every sender's packed sub-bucket histogram ((k-mers << 32) | records) of MY buckets, summed: the owner's scatter
        # offsets and the size of its k-mer list, without reading any record twice
        h_local = hist if torch.is_tensor(hist) else device_view(torch, hist, P * nb2, device)
        n_mine = (q_lo[rank + 1] - q_lo[rank]) * nb2
        h_recv = torch.empty(world * n_mine, dtype=torch.int64, device=device)
        dist.all_to_all_single(h_recv, h_local.contiguous(), output_split_sizes=[n_mine] * world,
                               input_split_sizes=[(q_lo[r + 1] - q_lo[r]) * nb2 for r in range(world)])
        staged_hist = h_recv.view(world, n_mine).sum(dim=0).contiguous()
        gathered_all = _to_host(ctx, torch.stack(gathered), torch, device)
        sizes_all = np.ascontiguousarray(gathered_all[:, :P])          # [sender, bucket], in records
        _stream_sync(torch, device)
        mark("sizes+sub_hist_exchange")
        seg_begin, seg_len = owner_segments(sizes_all, q_lo, rank)
        recv_records = int(seg_len.sum())
        if use_peer:
            all_handles = [np.ascontiguousarray(gathered_all[r, P:]).tobytes() for r in range(world)]
            peers = peer_pointers(ctx, torch, dist, device, handle, recs, cache_name="_pc_skm_peers", handles=all_handles)
            mark("peer_handles")
            seg_base = np.repeat(np.asarray(peers, dtype=np.uint64)[None, :], seg_len.shape[0], axis=0)
            ctx.skm_count_from(k, skm, seg_base, seg_begin, seg_len, staged_hist)
            exchanged = 16 * (recv_records - int(sizes_all[rank, q_lo[rank]:q_lo[rank + 1]].sum()))
        else:
            # all-to-all of the records themselves (two int64 per record), then count from the receive buffer
            send = recs if torch.is_tensor(recs) else device_view(torch, recs, 2 * n_local_records, device)
            send_counts = [2 * int(sizes_all[rank, q_lo[r]:q_lo[r + 1]].sum()) for r in range(world)]
            recv_counts = [2 * int(sizes_all[s_, q_lo[rank]:q_lo[rank + 1]].sum()) for s_ in range(world)]
            recv = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=device)
            dist.all_to_all_single(recv, send[:2 * n_local_records], output_split_sizes=recv_counts, input_split_sizes=send_counts)
            _stream_sync(torch, device)
            mark("record_all_to_all")
            mine = sizes_all[:, q_lo[rank]:q_lo[rank + 1]]
            recv_off = np.concatenate([[0], np.cumsum(recv_counts)])[:-1] // 2
            within = np.cumsum(mine, axis=1) - mine
            ctx.skm_count_from(k, skm, None, (recv_off[:, None] + within).T.copy(), mine.T.copy(), staged_hist, recs=recv)
            del recv
            exchanged = 8 * (int(sum(send_counts)) - send_counts[rank])
        n_local = int(ctx.count_info()["valid_windows"])
    else:
        kbuf_keys = int((ends - begins).sum()) if len(ends) else 1        # >= local windows: pins the bucket buffer
        if hasattr(ctx, "kbuf_export") and on_gpu and exchange != "nccl":
            ctx._pc_handle = ctx.kbuf_export(max(kbuf_keys, getattr(ctx, "_pc_kbuf_keys", 1)))
            ctx._pc_kbuf_keys = max(kbuf_keys, getattr(ctx, "_pc_kbuf_keys", 1))
        off, keys = ctx.partition(k, pb)
        mark("pack+partition")
        n_local = int(off[-1])
        sizes = torch.from_numpy(np.diff(off).astype(np.int64)).to(device)
        gathered = [torch.empty(P, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(gathered, sizes)
        sizes_all = _to_host(ctx, torch.stack(gathered), torch, device)   # [sender, bucket]
        send_counts = [int(sizes_all[rank, q_lo[r]:q_lo[r + 1]].sum()) for r in range(world)]
        recv_counts = [int(sizes_all[s_, q_lo[rank]:q_lo[rank + 1]].sum()) for s_ in range(world)]
        mine = sizes_all[:, q_lo[rank]:q_lo[rank + 1]]                    # [sender, my bucket]
        use_peer = exchange in ("peer", "kmers") and on_gpu or (exchange == "auto" and on_gpu
                                                                 and hasattr(ctx, "kbuf_export") and not torch.is_tensor(keys))
        mark("bucket_sizes_allgather")
        staged_hist = None
        if use_peer and hasattr(ctx, "sub_plan"):
            # shared-memory count: the owner needs the sizes of its sub-buckets before it scatters.  Every sender histograms
            # its own bucket buffer (local HBM) and the small histograms travel, instead of the owner streaming every peer
            # segment over NVLink a second time.  One fan-out for all ranks, from the all-gathered bucket sizes.
            nb2 = max(ctx.sub_plan(int(sizes_all[:, q_lo[r]:q_lo[r + 1]].sum()), max(1, q_lo[r + 1] - q_lo[r]))
                      for r in range(world))
            if nb2 > 0 and all(q_lo[r + 1] > q_lo[r] for r in range(world)):
                h_local = torch.empty(P * nb2, dtype=torch.int64, device=device)
                ctx.sub_hist(pb, nb2, h_local)
                n_mine = (q_lo[rank + 1] - q_lo[rank]) * nb2
                h_recv = torch.empty(world * n_mine, dtype=torch.int64, device=device)
                dist.all_to_all_single(h_recv, h_local, output_split_sizes=[n_mine] * world,
                                       input_split_sizes=[(q_lo[r + 1] - q_lo[r]) * nb2 for r in range(world)])
                staged_hist = h_recv.view(world, n_mine).sum(dim=0).contiguous()
                _stream_sync(torch, device)
                ctx.sub_hist_set(nb2, staged_hist)
                mark("sub_hist_exchange")
        if use_peer:
            # 2a. fused exchange: the owner's count kernel reads every sender's slice straight out of that sender's bucket
            # buffer over NVLink (CUDA IPC mapping).  The all-gather of the bucket sizes above is the barrier that says
            # "every rank has finished its partition"; the repeat-set all-gather below says "every rank has finished
            # reading", so the next step may overwrite the buffers.
            peers = peer_pointers(ctx, torch, dist, device, ctx._pc_handle, keys)
            mark("peer_handles")
            seg_begin, seg_len = owner_segments(sizes_all, q_lo, rank)
            seg_base = np.repeat(np.asarray(peers, dtype=np.uint64)[None, :], seg_len.shape[0], axis=0)
            ctx.count_buckets_from(k, pb, seg_base, seg_begin, seg_len)
            exchanged = 8 * (int(sum(recv_counts)) - recv_counts[rank])    # bytes read from peers' HBM over NVLink
        else:
            # 2b. all-to-all of the raw k-mers over NVLink (NCCL); count at the owner
            send = keys if torch.is_tensor(keys) else device_view(torch, keys, n_local, device)
            recv = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=device)
            dist.all_to_all_single(recv, send[:n_local], output_split_sizes=recv_counts, input_split_sizes=send_counts)
            mark("kmer_all_to_all")
            recv_off = np.concatenate([[0], np.cumsum(recv_counts)])[:-1]
            within = np.cumsum(mine, axis=1) - mine                           # offset of bucket q inside sender s's slice
            seg_begin = (recv_off[:, None] + within).T.copy()                 # [my bucket, sender]
            seg_len = mine.T.copy()
            ctx.count_buckets(k, pb, recv, seg_begin, seg_len)
            del recv
            exchanged = 8 * (n_local - send_counts[rank])
    if hasattr(ctx, "tune"):
        ctx.tune(keep_min=1)
    owner_info = ctx.count_info()
    info = dict(owner_info, valid_windows=n_local)
    mark("count_buckets")
    n_owner_sel = _select_raw(ctx, effective_low_count(k, kmer_low_count), kmer_high_count, bool(high_bool))
    mark("select")
    if hasattr(ctx, "timings"):
        tm = ctx.timings()
        phase.update({"select.k_select": tm.get("select", 0.0), "select.sort": tm.get("sort", 0.0),
                      "select.rep_table": tm.get("rep_table", 0.0)})
    sel_local = torch.empty(max(n_owner_sel, 1), dtype=torch.int64, device=device)
    if n_owner_sel:
        ctx.selected_into(sel_local)
    mark("selected_copy")
    # 3. all-gather the repeat set; every rank sorts it to the same column order
    sel_all, _ = _all_gather_v(dist, sel_local[:n_owner_sel], torch, ctx)
    if sampling_sensitivity > 1 and sel_all.numel():
        # sampling is defined on the globally sorted survivor list (pc_select semantics)
        sel_all = torch.sort(sel_all.view(torch.int64) ^ (-2**63))[0] ^ (-2**63)       # unsigned order
        sel_all = sel_all[::int(sampling_sensitivity)].contiguous()
    mark("selected_allgather")
    n_sel = ctx.set_selected(sel_all, k, n=sel_all.numel())
    # 4. local matrix, global idf
    nnz = ctx.build_matrix(local_chunk_row, n_local_rows)
    df_t = torch.zeros(max(n_sel, 1), dtype=torch.int32, device=device)
    if n_sel:
        ctx.df_into(df_t)
    mark("matrix")
    dist.all_reduce(df_t, op=dist.ReduceOp.SUM)
    mark("df_allreduce")
    if tfidf:
        ctx.tfidf(len(scaffolds_all), df_t if n_sel else None)
    mark("tfidf")
    t = torch.tensor([n_local, owner_info["distinct"], nnz], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    stats = dict(info)
    t = _to_host(ctx, t, torch, device)
    stats.update(n_selected=n_sel, nnz=nnz, n_rows=n_local_rows, n_rows_global=len(scaffolds_all), row0=row0,
                 valid_windows_global=int(t[0]), distinct_global=int(t[1]), nnz_global=int(t[2]),
                 owner_distinct=owner_info["distinct"], local_bases=int(ends[-1]) if len(ends) else 0,
                 phase_ms={k_: round(v, 2) for k_, v in phase.items()}, exchange_bytes_sent=exchanged,
                 exchange=("peer" if use_peer else "nccl_all_to_all") + ("/super-k-mers" if skm is not None else ""))
    scaffolds = scaffolds_all[row0:row0 + n_local_rows]
    if not fetch:
        return HotPathResult(layout, k, scaffolds, np.empty(0, np.uint64), None, None, None, None, None, stats,
                             ctx.timings())
    return fetch_result(ctx, layout, k, scaffolds, n_sel, nnz, tfidf, stats, arena, df=df_t[:n_sel].cpu().numpy())


def rerun_threshold(ctx, layout, ids, device, k=23, kmer_low_count=100, high_bool=0, kmer_high_count=2000000,
                    sampling_sensitivity=1, remove_non_chunk=1, min_chunk_threshold=0, min_chunk_size=0, prefilter=0,
                    tfidf=True, fetch=True, arena=None):
    """Threshold sweep (BASELINE config 4): the owners' count tables of a previous run_hot_path_distributed call on
    this context are kept; only K4 -> all-gather of the repeat set -> K5..K7 are repeated for the new threshold."""
    import torch
    import torch.distributed as dist
    bind_context_stream(ctx, torch, device)
    scaffolds_all, chunk_row_all = layout.rows(remove_non_chunk, min_chunk_threshold, min_chunk_size, prefilter)
    my_rows_global = chunk_row_all[ids]
    kept = np.flatnonzero(my_rows_global >= 0)
    row0 = int(my_rows_global[kept].min()) if len(kept) else 0
    local_chunk_row = np.where(my_rows_global >= 0, my_rows_global - row0, -1).astype(np.int32)
    n_local_rows = len(kept)
    n_owner_sel = _select_raw(ctx, effective_low_count(k, kmer_low_count), kmer_high_count, bool(high_bool))
    sel_local = torch.empty(max(n_owner_sel, 1), dtype=torch.int64, device=device)
    if n_owner_sel:
        ctx.selected_into(sel_local)
    sel_all, _ = _all_gather_v(dist, sel_local[:n_owner_sel], torch, ctx)
    if sampling_sensitivity > 1 and sel_all.numel():
        sel_all = torch.sort(sel_all.view(torch.int64) ^ (-2**63))[0] ^ (-2**63)
        sel_all = sel_all[::int(sampling_sensitivity)].contiguous()
    _stream_sync(torch, device)
    n_sel = ctx.set_selected(sel_all, k, n=sel_all.numel())
    nnz = ctx.build_matrix(local_chunk_row, n_local_rows)
    df_t = torch.zeros(max(n_sel, 1), dtype=torch.int32, device=device)
    if n_sel:
        ctx.df_into(df_t)
    dist.all_reduce(df_t, op=dist.ReduceOp.SUM)
    _stream_sync(torch, device)
    if tfidf:
        ctx.tfidf(len(scaffolds_all), df_t if n_sel else None)
    t = torch.tensor([nnz], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    stats = dict(n_selected=n_sel, nnz=nnz, n_rows=n_local_rows, n_rows_global=len(scaffolds_all), row0=row0,
                 nnz_global=int(t[0]))
    scaffolds = scaffolds_all[row0:row0 + n_local_rows]
    if not fetch:
        return HotPathResult(layout, k, scaffolds, np.empty(0, np.uint64), None, None, None, None, None, stats,
                             ctx.timings())
    return fetch_result(ctx, layout, k, scaffolds, n_sel, nnz, tfidf, stats, arena, df=df_t[:n_sel].cpu().numpy())


def gather_rows(result, dst=0):
    """Concatenate the shard-local CSR blocks on rank `dst` (host side, gather_object).  Returns the global
    HotPathResult on dst, None elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = (result.stats["row0"], result.scaffolds, result.indptr, result.indices, result.data, result.tfidf)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    gathered.sort(key=lambda p: (p[0], len(p[1]) == 0))
    scaffolds, indptr, indices, data, tf = [], [np.zeros(1, np.int64)], [], [], []
    base = 0
    for _row0, sc, ip, ix, dt, tv in gathered:
        scaffolds += sc
        indptr.append(ip[1:] + base)
        base += int(ip[-1])
        indices.append(ix); data.append(dt)
        if tv is not None:
            tf.append(tv)
    stats = dict(result.stats)
    stats.update(n_rows=len(scaffolds), nnz=base)
    return HotPathResult(result.layout, result.k, scaffolds, result.kmer_keys, np.concatenate(indptr),
                         np.concatenate(indices) if indices else np.zeros(0, np.int32),
                         np.concatenate(data) if data else np.zeros(0, np.float32),
                         np.concatenate(tf) if tf else None, result.df, stats, result.timings)
