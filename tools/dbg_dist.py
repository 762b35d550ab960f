import os, sys, socket
sys.path.insert(0, '/root/repo')
import numpy as np

def worker(rank, world, port, gsize, low):
    import torch, torch.distributed as dist
    from polycracker_b200 import binding, distributed as pcd, synth
    from polycracker_b200.pipeline import ChunkLayout, run_hot_path
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    p = synth.make_params(20260002, gsize)
    seq, off, names = synth.generate(p, threads=8)
    layout = ChunkLayout(names, off, 75000)
    def fetch(ids, begins, ends, total):
        buf = np.empty(total, np.uint8)
        for j, i in enumerate(ids.tolist()):
            buf[begins[j]:ends[j]] = seq[layout.chunk_begin_abs[i]:layout.chunk_end_abs[i]]
        return buf
    ctx = binding.Context(rank); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    res = pcd.run_hot_path_distributed(ctx, layout, fetch, dev, chunk_size=75000, k=23, kmer_low_count=low)
    print(rank, {k: res.stats[k] for k in ('n_selected','nnz','n_rows','valid_windows_global','distinct_global','nnz_global','owner_distinct','distinct')}, flush=True)
    dist.barrier()
    if rank == 0:
        ctx2 = binding.Context(0)
        r1 = run_hot_path(ctx2, seq, off, names, chunk_size=75000, k=23, kmer_low_count=low)
        print('single', r1.stats['n_selected'], r1.stats['nnz'], r1.stats['valid_windows'], r1.stats['distinct'], flush=True)
    dist.barrier(); dist.destroy_process_group()

if __name__ == '__main__':
    import torch.multiprocessing as mp
    gsize = int(sys.argv[1]); low = int(sys.argv[2])
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port, gsize, low), nprocs=2, join=True)
