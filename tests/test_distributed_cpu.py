"""CPU-only, world_size 2, gloo: the multi-GPU driver (sharding by sorted chunk blocks, owner exchange of the
counts, all-gather of the repeat set, df all-reduce, row-block gather).  The GPU context is replaced by a stand-in
backed by the C oracle, so what is under test is the host-side plumbing of polycracker_b200.distributed; the same
driver runs unchanged over NCCL with the real context (tests/test_gpu.py covers the device side of the exchange)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleContext:
    """Stand-in for binding.Context with the subset of methods the distributed driver calls (host tensors)."""

    def __init__(self):
        self.k = None

    def load_sequences(self, buf, begins, ends):
        self.seq = np.asarray(buf, dtype=np.uint8)
        self.begins, self.ends = np.asarray(begins, np.int64), np.asarray(ends, np.int64)

    def count(self, k, capacity=0):
        from oracle import c_oracle
        self.k = k
        self.keys, self.counts, self.n_windows = c_oracle.count(self.seq, self.begins, self.ends, k)

    def count_info(self):
        return dict(capacity=0, distinct=len(self.keys), valid_windows=int(getattr(self, "n_windows", 0)),
                    table_bytes=0, max_probe=0, k=self.k)

    def partition(self, k, pb):
        """Stand-in for pc_partition: every canonical k-mer INSTANCE of the local chunks, grouped by a hash bucket."""
        import torch
        from polycracker_b200 import binding
        words, nmask, _ = binding.host_pack_chunks(self.seq, self.begins, self.ends)
        keys, valid = binding.host_extract(words, nmask, k)
        inst = keys[valid]
        with np.errstate(over="ignore"):
            h = inst * np.uint64(0x9E3779B97F4A7C15)
        bucket = (h >> np.uint64(64 - pb)).astype(np.int64)
        order = np.argsort(bucket, kind="stable")
        self.k = k
        off = np.concatenate([[0], np.cumsum(np.bincount(bucket, minlength=1 << pb))]).astype(np.int64)
        return off, torch.from_numpy(inst[order].view(np.int64).copy())

    def count_buckets(self, k, pb, keys, seg_begin, seg_len, capacity=0):
        arr = keys.numpy().view(np.uint64)
        parts = [arr[b:b + n] for b, n in zip(np.asarray(seg_begin).ravel().tolist(), np.asarray(seg_len).ravel().tolist())]
        allk = np.concatenate(parts) if parts else np.empty(0, np.uint64)
        self.k = k
        self.keys, cnt = np.unique(allk, return_counts=True)
        self.counts = cnt.astype(np.uint64)
        self.n_windows = 0

    def export_by_owner(self, world, keys_t, counts_t):
        owner = (self.keys % np.uint64(world)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        keys_t.numpy()[:len(order)] = self.keys[order].view(np.int64)
        counts_t.numpy()[:len(order)] = self.counts[order].astype(np.int32)
        return np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=world))]).astype(np.int64)

    def merge_counts(self, keys_t, counts_t, n, k, capacity=0, fresh=True):
        keys = keys_t.numpy()[:n].view(np.uint64)
        cnt = counts_t.numpy()[:n].astype(np.uint64)
        self.k = k
        self.keys, inv = np.unique(keys, return_inverse=True)
        self.counts = np.bincount(inv, weights=cnt, minlength=len(self.keys)).astype(np.uint64)
        self.n_windows = 0

    def select(self, low, high=2000000, use_high=False, sampling=1):
        keep = self.counts >= low
        if use_high:
            keep &= self.counts <= high
        self.sel = self.keys[keep][::max(1, sampling)]
        return len(self.sel)

    def selected_into(self, t):
        t.numpy()[:len(self.sel)] = self.sel.view(np.int64)

    def selected(self, n):
        return self.sel[:n].copy()

    def set_selected(self, keys_t, k, n=None):
        self.sel = np.unique(keys_t.numpy()[:n].view(np.uint64))
        return len(self.sel)

    def build_matrix(self, chunk_row, n_rows):
        from oracle import c_oracle
        chunk_row = np.asarray(chunk_row)
        row_chunk = np.full(n_rows, -1, np.int64)
        for c, r in enumerate(chunk_row.tolist()):
            if r >= 0:
                row_chunk[r] = c
        self.indptr, self.indices, self.data = c_oracle.matrix(self.seq, self.begins, self.ends, row_chunk, self.k, self.sel)
        return len(self.data)

    def df_into(self, t):
        t.numpy()[:len(self.sel)] = np.bincount(self.indices, minlength=len(self.sel)).astype(np.int32)

    def tfidf(self, n_rows_global=0, df_global=None):
        df = df_global.numpy()[:len(self.sel)].astype(np.float32)
        idf = (np.log(np.float32(n_rows_global + 1) / (df + np.float32(1))) + np.float32(1)).astype(np.float32)
        x = (self.data * idf[self.indices]).astype(np.float32)
        out = x.copy()
        for r in range(len(self.indptr) - 1):
            seg = x[self.indptr[r]:self.indptr[r + 1]]
            s = np.sqrt(np.sum((seg * seg).astype(np.float64)))
            if s > 0:
                out[self.indptr[r]:self.indptr[r + 1]] = (seg.astype(np.float64) / s).astype(np.float32)
        self.tf = out

    def csr(self, n_rows, nnz):
        return self.indptr, self.indices, self.data

    def tfidf_values(self, nnz):
        return self.tf

    def timings(self):
        return {}


class SkmOracleContext(OracleContext):
    """Adds the super-k-mer entry points (pc_skm_*): the records come from the device's own extraction routines run on the
    host (binding.host_skm_records), bucketed like pc_skm_partition; the owner expands and counts what it received."""

    def skm_plan(self, k, windows_global, world):
        from polycracker_b200 import binding
        pb, nb2 = 2, 3
        m = binding.host_skm_m(k, (1 << pb) * nb2)
        return np.array([pb, nb2, m, k - m + 1, min(16, 49 - k + 1), 1 if m else 0, 0, 0], np.int32)

    def skm_partition(self, k, plan):
        import torch
        from polycracker_b200 import binding
        pb, nb2, m, nmax = int(plan[0]), int(plan[1]), int(plan[2]), int(plan[4])
        words, nmask, _ = binding.host_pack_chunks(self.seq, self.begins, self.ends)
        recs = binding.host_skm_records(words, nmask, k, m, nmax, pb, nb2)
        dest = ((recs[:, 1] >> np.uint64(6)) & np.uint64(0xFFFFFF)).astype(np.int64)
        q, sub = dest >> 11, dest & 2047
        n = (recs[:, 1] & np.uint64(63)).astype(np.int64) + 1
        order = np.argsort(q, kind="stable")
        hist = np.zeros((1 << pb) * nb2, np.int64)
        np.add.at(hist, q * nb2 + sub, (n << 32) | 1)
        self.k, self.n_windows = k, int(n.sum())
        off = np.concatenate([[0], np.cumsum(np.bincount(q, minlength=1 << pb))]).astype(np.int64)
        return off, torch.from_numpy(recs[order].reshape(-1).view(np.int64).copy()), torch.from_numpy(hist)

    def skm_count_from(self, k, plan, seg_base, seg_begin, seg_len, hist, recs=None):
        from polycracker_b200 import binding
        assert seg_base is None and recs is not None            # CPU: the all-to-all variant of the exchange
        arr = recs.numpy().view(np.uint64).reshape(-1, 2)
        parts = [arr[b:b + n] for b, n in zip(np.asarray(seg_begin).ravel().tolist(), np.asarray(seg_len).ravel().tolist())]
        mine = np.concatenate(parts) if parts else np.empty((0, 2), np.uint64)
        h = hist.numpy()
        assert int((h & 0xFFFFFFFF).sum()) == len(mine)         # the summed histogram describes exactly what arrived
        keys, _ = binding.host_skm_expand(mine, k)
        assert int((h >> 32).sum()) == len(keys)
        self.keys, cnt = np.unique(keys, return_counts=True)
        self.counts = cnt.astype(np.uint64)


def _worker(rank, world, port, tmpdir, sampling, skm=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from polycracker_b200 import distributed as pcd, synth
    from polycracker_b200.pipeline import ChunkLayout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = synth.small_params(seed=91, genome_size=600_000, n_chrom=3)
        lay_all = synth.layout(p)
        names = [n for n, _ in lay_all]
        offs = np.concatenate([[0], np.cumsum([l for _, l in lay_all])]).astype(np.int64)
        layout = ChunkLayout(names, offs, 20000)

        def fetch(ids, begins, ends, total):
            buf = np.empty(total, np.uint8)
            for c in sorted(set(layout.chunk_seq[ids].tolist())):        # a rank only makes what it needs
                chrom, _, _ = synth.generate(p, chroms=[c])
                for j in np.flatnonzero(layout.chunk_seq[ids] == c).tolist():
                    i = ids[j]
                    buf[begins[j]:ends[j]] = chrom[layout.chunk_start[i]:layout.chunk_end[i]]
            return buf

        res = pcd.run_hot_path_distributed(SkmOracleContext() if skm else OracleContext(), layout, fetch,
                                           torch.device("cpu"), chunk_size=20000, k=23, kmer_low_count=10,
                                           sampling_sensitivity=sampling)
        assert res.stats["exchange"] == ("nccl_all_to_all/super-k-mers" if skm else "nccl_all_to_all")
        full = pcd.gather_rows(res)
        if rank == 0:
            np.savez(os.path.join(tmpdir, "out.npz"), keys=full.kmer_keys, indptr=full.indptr, indices=full.indices,
                     data=full.data, tfidf=full.tfidf, scaffolds=np.array(full.scaffolds),
                     stats=np.array([full.stats["valid_windows_global"], full.stats["distinct_global"],
                                     full.stats["nnz_global"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sampling,skm", [(1, False), (3, False), (1, True)])
def test_two_rank_gloo_matches_single_process_oracle(tmp_path, sampling, skm):
    import torch.multiprocessing as mp
    from polycracker_b200 import synth
    from tests import util
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), sampling, skm), nprocs=2, join=True)
    got = np.load(tmp_path / "out.npz")
    p = synth.small_params(seed=91, genome_size=600_000, n_chrom=3)
    seq, off, names = synth.generate(p)
    want = util.c_oracle_path(seq, off, names, 20000, 23, 10, sampling=sampling)
    assert got["keys"].tolist() == want["sel"].tolist()
    assert got["scaffolds"].tolist() == want["scaffolds"]
    assert np.array_equal(got["indptr"], want["indptr"]) and np.array_equal(got["indices"], want["indices"])
    assert np.array_equal(got["data"], want["data"])
    util.assert_tfidf_close(got["tfidf"], want["tfidf"])
    assert got["stats"].tolist() == [want["n_windows"], len(want["keys"]), len(want["data"])]
    assert len(want["sel"]) > 10 and len(want["data"]) > 100


def test_shard_chunks_blocks_are_contiguous_and_balanced():
    from polycracker_b200 import distributed as pcd
    from polycracker_b200.pipeline import ChunkLayout
    names = ["a", "a_1", "chr10", "chr1", "b"]                # names whose chunk labels interleave when sorted
    lens = [95, 40, 300, 120, 7]
    offs = np.concatenate([[0], np.cumsum(lens)])
    lay = ChunkLayout(names, offs, 10)
    order = sorted(range(lay.n_chunks), key=lambda i: lay.chunk_names[i])
    for world in (1, 2, 3, 8):
        blocks = pcd.shard_chunks(lay, world)
        assert np.concatenate(blocks).tolist() == order       # a partition of the sorted list, in order
        sizes = [int((lay.chunk_end[b] - lay.chunk_start[b]).sum()) for b in blocks]
        assert max(sizes) - min(sizes) <= 2 * 10 + 10
        for b in blocks:
            ids, begins, ends, total = pcd.local_buffer_plan(lay, b)
            assert total == int((lay.chunk_end[ids] - lay.chunk_start[ids]).sum())
            assert np.all(begins[1:] == ends[:-1])
