"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the real contexts over NCCL, one process per GPU, against the
single-process C oracle.  The same driver is covered on CPU (gloo, world 2) by tests/test_distributed_cpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir, part_mode, exchange):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from polycracker_b200 import binding, distributed as pcd, synth
    from polycracker_b200.pipeline import ChunkLayout
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        p = synth.small_params(seed=191, genome_size=3_000_000, n_chrom=3, n_shared_families=6, n_specific_families=8)
        lay_all = synth.layout(p)
        names = [n for n, _ in lay_all]
        offs = np.concatenate([[0], np.cumsum([l for _, l in lay_all])]).astype(np.int64)
        layout = ChunkLayout(names, offs, 20000)

        def fetch(ids, begins, ends, total):
            buf = np.empty(total, np.uint8)
            for c in sorted(set(layout.chunk_seq[ids].tolist())):
                chrom, _, _ = synth.generate(p, chroms=[c])
                for j in np.flatnonzero(layout.chunk_seq[ids] == c).tolist():
                    i = ids[j]
                    buf[begins[j]:ends[j]] = chrom[layout.chunk_start[i]:layout.chunk_end[i]]
            return buf

        ctx = binding.Context(rank)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.tune(part_mode=part_mode, part_mbytes=1)
        for _ in range(2):                                    # twice: the peer mappings and buffers are re-used
            res = pcd.run_hot_path_distributed(ctx, layout, fetch, dev, chunk_size=20000, k=23, kmer_low_count=10,
                                               exchange=exchange)
        skm = part_mode in (3, -1)                            # k = 23: the automatic mode takes the super-k-mer path
        assert res.stats["exchange"] == ("nccl_all_to_all" if exchange == "nccl" else "peer") + ("/super-k-mers" if skm else "")
        if exchange == "peer":                                # threshold sweep on the kept count tables (config 4)
            blocks = pcd.shard_chunks(layout, world)
            ids = pcd.local_buffer_plan(layout, blocks[rank])[0]
            res2 = pcd.rerun_threshold(ctx, layout, ids, dev, k=23, kmer_low_count=25)
            full2 = pcd.gather_rows(res2)
            if rank == 0:
                np.savez(os.path.join(tmpdir, "out25.npz"), keys=full2.kmer_keys, indptr=full2.indptr,
                         indices=full2.indices, data=full2.data, tfidf=full2.tfidf)
        full = pcd.gather_rows(res)
        if rank == 0:
            np.savez(os.path.join(tmpdir, "out.npz"), keys=full.kmer_keys, indptr=full.indptr, indices=full.indices,
                     data=full.data, tfidf=full.tfidf, scaffolds=np.array(full.scaffolds),
                     stats=np.array([full.stats["valid_windows_global"], full.stats["distinct_global"],
                                     full.stats["nnz_global"]]))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("part_mode,exchange", [(0, "nccl"), (1, "nccl"), (1, "peer"), (2, "nccl"), (2, "peer"), (-1, "peer"),
                                                (3, "peer"), (3, "nccl"), (-1, "auto")])
def test_two_gpus_nccl_match_oracle(tmp_path, part_mode, exchange):
    from polycracker_b200 import binding, synth
    from tests import util
    if binding.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), part_mode, exchange), nprocs=2, join=True)
    got = np.load(tmp_path / "out.npz")
    p = synth.small_params(seed=191, genome_size=3_000_000, n_chrom=3, n_shared_families=6, n_specific_families=8)
    seq, off, names = synth.generate(p)
    want = util.c_oracle_path(seq, off, names, 20000, 23, 10)
    assert got["keys"].tolist() == want["sel"].tolist()
    assert got["scaffolds"].tolist() == want["scaffolds"]
    assert np.array_equal(got["indptr"], want["indptr"]) and np.array_equal(got["indices"], want["indices"])
    assert np.array_equal(got["data"], want["data"])
    util.assert_tfidf_close(got["tfidf"], want["tfidf"])
    assert got["stats"].tolist() == [want["n_windows"], len(want["keys"]), len(want["data"])]
    assert len(want["sel"]) > 100
    if exchange == "peer":
        got25 = np.load(tmp_path / "out25.npz")
        want25 = util.c_oracle_path(seq, off, names, 20000, 23, 25)
        assert got25["keys"].tolist() == want25["sel"].tolist() and len(want25["sel"]) < len(want["sel"])
        assert np.array_equal(got25["indptr"], want25["indptr"]) and np.array_equal(got25["indices"], want25["indices"])
        assert np.array_equal(got25["data"], want25["data"])
        util.assert_tfidf_close(got25["tfidf"], want25["tfidf"])
