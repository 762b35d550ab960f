#!/usr/bin/env python
"""Do N ranks get N x the host<->device bandwidth of one?  (torchrun, one rank per GPU.)  Each rank copies a 512 MB pinned
buffer to its GPU (and back) ten times, all ranks at once, for several ways of providing the pinned memory:
  torch      torch.empty(pin_memory=True)                     (what bench.py uses)
  thp        numpy buffer with madvise(MADV_HUGEPAGE), touched, then cudaHostRegister
  both       H2D and D2H at the same time (what the pipelined e2e loop does)
Prints per-rank GB/s (max-over-ranks time) and the aggregate."""
import ctypes, mmap, os, sys, time
import numpy as np, torch, torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 512 << 20
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
cudart = torch.cuda.cudart()

def thp_pinned(nbytes):
    m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    try:
        m.madvise(mmap.MADV_HUGEPAGE)
    except Exception as e:
        print("madvise failed", e)
    a = np.frombuffer(m, dtype=np.uint8)
    a[:] = 1                                                   # touch (first touch after the advice)
    t = torch.from_numpy(a)
    r = cudart.cudaHostRegister(t.data_ptr(), nbytes, 0)
    assert int(r) == 0, r
    return t, m

def run(name, h_in, h_out, both=False):
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def once(which):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            if which in ("h2d", "both"):
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if which in ("d2h", "both"):
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    once("h2d")
    res = {}
    for which in ("h2d", "d2h", "both"):
        dt = once(which)
        res[which] = 10 * n * (2 if which == "both" else 1) / dt / 1e9
    if rank == 0:
        print("%-6s N=%d per-rank GB/s: h2d %.1f  d2h %.1f  both %.1f | aggregate: h2d %.1f  d2h %.1f  both %.1f" % (
            name, world, res["h2d"], res["d2h"], res["both"], world * res["h2d"], world * res["d2h"], world * res["both"]), flush=True)

a = torch.empty(n, dtype=torch.uint8, pin_memory=True); b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
a.fill_(1)
run("torch", a, b)
try:
    c, mc = thp_pinned(n); d, md = thp_pinned(n)
    if rank == 0:
        print("AnonHugePages:", [l.strip() for l in open("/proc/meminfo") if "AnonHugePages" in l])
    run("thp", c, d)
except Exception as e:
    if rank == 0: print("thp variant failed:", repr(e))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
