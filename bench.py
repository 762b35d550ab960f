#!/usr/bin/env python
"""bench.py -- matrix-build Gbp/s at k=23 (BASELINE.json metric) on 1..8 B200, next to the host-CPU reference.

    python bench.py --gpus N --steps K --warmup W              # the CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W      # the CPU restatement of the reference path

A "step" is one pass of the hot path (2-bit pack -> count -> select -> exact match -> CSR -> TF-IDF) over the
synthetic genome of BASELINE.json configs[1] (500 Mbp allotetraploid, k=23, 75 kb chunks, threshold 30); with N GPUs
every rank gets one 500 Mbp share of an N x 500 Mbp genome of the same shape (weak scaling) and the count exchange
(all-to-all), repeat-set all-gather and df all-reduce run inside the step.

One JSON line on stdout (rank 0).  `value` = genome bases / step time with the ASCII genome already resident in HBM;
`e2e` = K back-to-back steps through the public API with pinned HOST buffers, every step's H2D of the genome and D2H of
the repeat set + CSR + TF-IDF inside the timed region, copies of neighbouring steps overlapping this step's kernels
(`e2e.sequential_*`: nothing overlapped); `roofline` is for the longest-running kernel from its CUDA-event time, with the
per-kernel table beside it; `cpu_baseline` is the C oracle (oracle/oracle.c, a port: the reference's own tool chain
cannot run) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm (rank 0 only) is meant to use every host core
    # it can.  libgomp reads the variable when it is first loaded, so set it before anything links it in.
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "matrix_build_gbp_per_s_k23"
UNIT = "Gbp/s"
WORKLOAD = "cfg2_500Mbp_allotetraploid"
FALLBACK_HBM_GBS = 6650.0
# dram__bytes_read.sum + dram__bytes_write.sum per step of the dominant unit (K3 = every kernel of the count), summed over
# its launches, from the committed `ncu --set full` capture of THIS workload with THIS kernel variant
# (profiles/r02_ncu_k3_500Mbp.md).  Keyed by (count kind, super-k-mer minimizer length); anything else reports null.
NCU_TRAFFIC = {
    # (unit, count kind, minimizer length, genome bases): 1.56 (k_skm_extract) + 2.81 (k_rec_scatter<1>) + 2.81 (k_rec_scatter<2>)
    # + 1.51 (k_count_skm) GB, profiles/r02_ncu_summary.md section 2
    ("K3_count", "list", 12, 497057026): 8.692e9,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-size", type=int, default=0, help="override bases per GPU (testing only)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="bases of the genome the CPU legs run on (0 = reference arm: the "
                    "whole genome; cpu_baseline leg of the GPU arm: a bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the 3 Mbp oracle check that precedes the timed loops")
    ap.add_argument("--no-iso", action="store_true", help="skip the iso-work (threshold x N) variant at N > 1")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe), sampled in-process through
    NVML (nvidia_ml_py) every 100 ms.  An `nvidia-smi -lms` child was measured to stall this process's CUDA calls for
    ~35 ms per query on these boxes (every other step of the timed loop took 70 ms instead of 36), so nvidia-smi is
    only the fallback, at a 1 s period.  Samples are time-stamped; only those inside [mark_begin, mark_end] count."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []                 # (t, sm_mhz, max_mhz, set(reasons))
        self.t0 = self.t1 = None
        self._stop = False
        self.proc = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        bits = get_reasons(h)
                        self.samples.append((time.perf_counter(), float(sm), float(mx),
                                             {n for n, b in names.items() if bits & b}))
                    except Exception:
                        pass
                    time.sleep(0.2)
            threading.Thread(target=loop, daemon=True).start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "1000"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            rs = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                  if v.lower().startswith("active")}
            self.samples.append((time.perf_counter(), sm, mx, rs))

    def wait_ready(self, timeout=20.0):
        t = time.perf_counter()
        while self.how and not self.samples and time.perf_counter() - t < timeout:
            time.sleep(0.05)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop = True
        if self.proc:
            self.proc.terminate()
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        inside = [x for x in self.samples if self.t0 is None or (self.t0 - 0.11 <= x[0] <= (self.t1 or x[0]) + 0.11)]
        use = inside or self.samples[-3:]
        reasons = set()
        for x in use:
            reasons |= x[3]
        return {"sm_mhz": float(np.median([x[1] for x in use])) if use else None,
                "sm_max_mhz": max([x[2] for x in use]) if use else None, "samples": len(use),
                "samples_in_timed_region": len(inside), "source": self.how, "reasons": sorted(reasons)}


def wait_for_quiet_gpu(gpu_index, timeout=20.0):
    """Timing hygiene: a process that has just exited (the test-suite, a previous bench) is still tearing its CUDA
    context down for a second or two and its driver locks stall this process's launches and synchronisations.  Wait
    until no other compute process is listed on the GPU (bounded)."""
    me = str(os.getpid())
    t0 = time.perf_counter()
    waited = 0.0
    while time.perf_counter() - t0 < timeout:
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-compute-apps=pid", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout
        except Exception:
            break
        others = [p.strip() for p in out.split("\n") if p.strip() and p.strip() != me]
        if not others:
            break
        time.sleep(0.5)
        waited = time.perf_counter() - t0
    return waited


def workload_params(n_gpus, genome_size_override):
    from polycracker_b200 import synth
    cfg = dict(synth.CONFIGS[WORKLOAD])
    per_gpu = genome_size_override or cfg["genome_size"]
    params = synth.make_params(cfg["seed"], per_gpu * n_gpus, n_subgenomes=cfg["n_subgenomes"])
    return cfg, params, per_gpu


def algorithmic_bytes_count(n_bases_packed, n_windows):
    """0.375 B per packed base (2-bit codes + 1-bit mask) + 24 B per k-mer instance (one 12-byte
    slot read-modify-write)."""
    return 0.375 * n_bases_packed + 24.0 * n_windows


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_pass(seq, off, names, cfg):
    """The C restatement of the reference path (oracle/ref_path.py -> oracle.c) over one in-memory genome; seconds."""
    from oracle import ref_path
    t0 = time.perf_counter()
    out = ref_path.run(seq, off, names, cfg["chunk_size"], cfg["k"], cfg["kmer_low_count"])
    return time.perf_counter() - t0, out


def run_reference_arm(args):
    """The reference's own path restated on the CPU, on the SAME genome and config as the GPU arm (the full cfg2 genome:
    every step is one whole pass), all host cores.  Nothing of the product is imported or mapped here: the genome comes
    from oracle/libpcsynth.so, geometry and row order from the literal oracle, the arithmetic from oracle.c."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle, ref_path
    cfg, params = ref_path.workload(WORKLOAD, args.genome_size)
    seq, off, names = ref_path.generate(params, max_bases=args.cpu_sample, threads=8)
    times = []
    for i in range(args.warmup + args.steps):
        dt, out = cpu_reference_pass(seq, off, names, cfg)
        if i >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    val = len(seq) / sec / 1e9
    cores = c_oracle.threads()
    sample = "%d bp = %s of the %s genome, full path count->select->match->CSR->TF-IDF" % (
        len(seq), "all %d chromosomes" % len(names) if not args.cpu_sample else "the first %d chromosomes" % len(names), WORKLOAD)
    mapped = [l.split()[-1] for l in open("/proc/self/maps") if ".so" in l and ("polycracker" in l or "oracle" in l)]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "genome_bases": int(len(seq)), "bases_per_gpu": int(len(seq)), "k": cfg["k"],
                       "chunk_size": cfg["chunk_size"], "kmer_low_count": cfg["kmer_low_count"],
                       "n_subgenomes": cfg["n_subgenomes"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "result": {"n_selected": int(len(out["sel"])), "nnz": int(len(out["data"])), "rows": int(len(out["scaffolds"])),
                       "valid_windows": int(out["n_windows"]), "distinct_kmers": int(len(out["keys"]))},
            "native_libraries_mapped": sorted(set(os.path.basename(m) for m in mapped)),
            "note": "oracle/oracle.c (C+OpenMP restatement): the reference's kmercountexact/bbmap/bedtools/py2 chain "
                    "cannot run here (SURVEY.md F8)"}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def parity_check(ctx, device, world, rank, dist):
    """Before anything is timed: a 3 Mbp genome through the SAME entry points the timed steps use (run_hot_path, or
    run_hot_path_distributed + gather_rows over NCCL at N > 1) against the CPU oracle -- repeat set, CSR bit-exact, TF-IDF
    within 1e-6 relative.  A bench line is only printed when this holds."""
    from polycracker_b200 import synth
    from polycracker_b200.pipeline import ChunkLayout, run_hot_path
    p = synth.small_params(seed=191, genome_size=3_000_000, n_chrom=3, n_shared_families=6, n_specific_families=8)
    k, chunk, low = 23, 20000, 10
    lay_all = synth.layout(p)
    names = [n for n, _ in lay_all]
    offs = np.concatenate([[0], np.cumsum([l for _, l in lay_all])]).astype(np.int64)
    layout = ChunkLayout(names, offs, chunk)
    seq, _, _ = synth.generate(p)
    if world == 1:
        full = run_hot_path(ctx, seq, offs, names, chunk_size=chunk, k=k, kmer_low_count=low, layout=layout)
    else:
        from polycracker_b200 import distributed as pcd

        def fetch(ids, begins, ends, total):
            buf = np.empty(total, np.uint8)
            for j, i in enumerate(ids.tolist()):
                a = int(offs[layout.chunk_seq[i]])
                buf[begins[j]:ends[j]] = seq[a + layout.chunk_start[i]:a + layout.chunk_end[i]]
            return buf
        res = pcd.run_hot_path_distributed(ctx, layout, fetch, device, chunk_size=chunk, k=k, kmer_low_count=low,
                                           exchange=os.environ.get("PC_EXCHANGE", "auto"))
        full = pcd.gather_rows(res)
    out = None
    if rank == 0:
        from oracle import ref_path
        want = ref_path.run(seq, offs, names, chunk, k, low)
        tf_err = 0.0
        if len(want["tfidf"]):
            tf_err = float(np.max(np.abs(np.asarray(full.tfidf, np.float64) - want["tfidf"]) / np.maximum(np.abs(want["tfidf"]), 1e-30)))
        ok = (np.array_equal(full.kmer_keys, want["sel"]) and full.scaffolds == want["scaffolds"]
              and np.array_equal(full.indptr, want["indptr"]) and np.array_equal(full.indices, want["indices"])
              and np.array_equal(full.data, want["data"]) and tf_err <= 1e-6)
        out = {"n_gpus": world, "ok": bool(ok), "genome_bases": int(len(seq)), "k": k, "n_selected": int(len(want["sel"])),
               "nnz": int(len(want["data"])), "rows": int(len(want["scaffolds"])), "tfidf_max_rel_err": tf_err,
               "checker": "oracle/ref_path.py (oracle.c)", "entry": "run_hot_path" if world == 1 else
               "run_hot_path_distributed + gather_rows (%s)" % full.stats.get("exchange")}
    if dist is not None:
        flag = [out["ok"] if rank == 0 else None]
        dist.broadcast_object_list(flag, src=0)
        if not flag[0]:
            raise SystemExit("parity check failed: %r" % (out,))
    elif not out["ok"]:
        raise SystemExit("parity check failed: %r" % (out,))
    return out


def run_b200_arm(args):
    import torch
    from polycracker_b200 import binding, synth
    from polycracker_b200.pipeline import ChunkLayout, HostArena, run_hot_path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available() or binding.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: polycracker_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    numa = None
    if world > 1:
        import torch.distributed as dist
        from polycracker_b200.distributed import bind_to_gpu_numa
        if os.environ.get("PC_NUMA_BIND", "1") != "0":
            numa = bind_to_gpu_numa(local_rank)             # before any pinned buffer is allocated (first touch)
        dist.init_process_group("nccl", device_id=device)

    cfg, params, per_gpu = workload_params(world, args.genome_size)
    k, chunk, low = cfg["k"], cfg["chunk_size"], cfg["kmer_low_count"]
    lay_all = synth.layout(params)
    names_all = [n for n, _ in lay_all]
    offs_all = np.concatenate([[0], np.cumsum([l for _, l in lay_all])]).astype(np.int64)
    genome_bases = int(offs_all[-1])
    layout = ChunkLayout(names_all, offs_all, chunk)

    quiet_wait = wait_for_quiet_gpu(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                    # long before the timed region, see ClockSampler
    ctx = binding.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- stage this rank's share of the genome in pinned host memory -----------------------------------------
    if world == 1:
        pinned = torch.empty(genome_bases, dtype=torch.uint8, pin_memory=True)
        synth.generate(params, out=pinned.numpy(), threads=8)
        ids = np.arange(layout.n_chunks)
        begins, ends = layout.chunk_begin_abs, layout.chunk_end_abs
        local_bases = genome_bases
    else:
        from polycracker_b200 import distributed as pcd
        blocks = pcd.shard_chunks(layout, world)
        ids, begins, ends, local_bases = pcd.local_buffer_plan(layout, blocks[rank])
        pinned = torch.empty(max(local_bases, 1), dtype=torch.uint8, pin_memory=True)
        host = pinned.numpy()
        need = sorted(set(layout.chunk_seq[ids].tolist()))
        for c in need:
            chrom, _, _ = synth.generate(params, chroms=[c], threads=4)
            for j in np.flatnonzero(layout.chunk_seq[ids] == c).tolist():
                i = ids[j]
                host[begins[j]:ends[j]] = chrom[layout.chunk_start[i]:layout.chunk_end[i]]
            del chrom
    # device-resident copy for the `value` loop (staged outside the timed region); K1 reads it in place
    dev_genome = pinned[:max(local_bases, 1)].to(device, non_blocking=False)
    arena = HostArena()

    def step(resident, fetch):
        src = dev_genome[:local_bases] if resident else pinned[:local_bases]
        if world == 1:
            return run_hot_path(ctx, src, offs_all, names_all, chunk_size=chunk, k=k, kmer_low_count=low,
                                fetch=fetch, layout=layout, arena=arena)
        return pcd.run_hot_path_distributed(ctx, layout, None, device, chunk_size=chunk, k=k, kmer_low_count=low,
                                            fetch=fetch, staged=(src, ids, begins, ends), arena=arena,
                                            exchange=os.environ.get("PC_EXCHANGE", "auto"))

    layout.rows()                                          # host-side row order: cached, reused by every step
    parity = None if args.no_parity else parity_check(ctx, device, world, rank, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step_wall = []
    # Host hygiene: a generation-2 garbage collection walks every object torch and numpy created at import
    # (30-100 ms on this image) and would land inside random steps of the timed loop.  Freeze what exists and keep the
    # collector off while timing; the steps themselves allocate no cycles.
    import gc
    gc.collect()
    gc.freeze()

    def timed(n_steps, resident, fetch, collect=None):
        gc.disable()
        try:
            return _timed(n_steps, resident, fetch, collect)
        finally:
            gc.enable()

    def _timed(n_steps, resident, fetch, collect=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n_steps):
            t_s = time.perf_counter()
            last = step(resident, fetch)
            if collect is not None:
                collect.append(ctx.timings())
                step_wall.append((time.perf_counter() - t_s) * 1e3)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------
    # The first dozen steps of a process still pay lazy driver / allocator work (sporadic +1.4 ms steps, see
    # tools/timeline.py): warm up for at least 12 steps, whatever smaller W was asked for (reported as `warmup_done`).
    warmup_done = max(args.warmup, 12)
    timed(warmup_done, True, False)
    ctx.reset_counters()
    if rank == 0:
        sampler.wait_ready()
        sampler.mark_begin()
    # A timed region of K steps that contains a host stall (another tenant of the box, a driver lock: single steps of
    # 200+ ms between 17 ms ones were seen on one pool box, kernels unaffected) measures the box, not the path.  The region
    # is therefore re-measured (whole K-step regions, at most 3) while one of its steps took more than twice the median
    # step; every attempt is reported (`timed_attempts`), the FIRST clean one -- or, if none is clean, the fastest -- is
    # the line's value.  All ranks take the same decision (the flag is all-reduced).
    attempts = []
    for attempt in range(3):
        per_step_timings = []
        n_wall = len(step_wall)
        ms_total, last_resident = timed(args.steps, True, False, per_step_timings)
        walls = step_wall[n_wall:]
        stalled = bool(walls) and max(walls) > 2.0 * float(np.median(walls)) and len(walls) >= 3
        if dist is not None:
            f = torch.tensor([1.0 if stalled else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(f, op=dist.ReduceOp.MAX)
            stalled = bool(f.item() > 0)
        attempts.append({"ms_per_step": ms_total / args.steps, "stalled_step": stalled, "timings": per_step_timings,
                         "last": last_resident, "walls": walls})
        if not stalled:
            break
    clean = [a for a in attempts if not a["stalled_step"]]
    chosen = clean[0] if clean else min(attempts, key=lambda a: a["ms_per_step"])
    ms_total, last_resident, per_step_timings = chosen["ms_per_step"] * args.steps, chosen["last"], chosen["timings"]
    step_wall = chosen["walls"]
    if rank == 0:
        sampler.mark_end()
    launches = ctx.launch_count()
    ms_per_step = ms_total / args.steps
    value = genome_bases / (ms_per_step * 1e-3) / 1e9
    info = ctx.count_info()
    detail_main = ctx.count_detail()
    main_phase = last_resident.stats.get("phase_ms")

    # ---- iso-work variant of the weak-scaling step (N > 1): with a fixed threshold the repeat set grows with the genome
    # (0.67 M k-mers at N = 1, 5.6 M at N = 8) and with it the matrix per rank, so part of the "scaling loss" is more work
    # per rank.  Here the threshold scales with N, which keeps the repeat set and the entries per rank about constant.
    iso = None
    if world > 1 and not args.no_iso:
        low_main = low
        low = low_main * world
        timed(2, True, False)
        n_iso = max(3, args.steps // 2)
        ms_iso, last_iso = timed(n_iso, True, False)
        iso = {"kmer_low_count": low, "steps": n_iso, "ms_per_step": ms_iso / n_iso,
               "value": genome_bases / (ms_iso / n_iso * 1e-3) / 1e9, "unit": UNIT,
               "n_selected": last_iso.stats.get("n_selected"), "nnz_rank0": last_iso.stats.get("nnz"),
               "phase_ms_rank0": last_iso.stats.get("phase_ms"),
               "note": "same genome and ranks, repeat-count threshold scaled by N (iso-work per rank); the headline `value` "
                       "keeps the fixed threshold of the named config"}
        low = low_main
        timed(1, True, False)                              # back to the named config for everything below

    # ---- end to end through the public API with host buffers (`e2e`) -------------------------------------------
    e2e = None
    res = None
    if not args.no_e2e:
        # (1) one isolated step, copies and kernels strictly in sequence: the latency a single call sees
        timed(1, False, True)
        ms_lat, res = timed(args.steps, False, True)
        # (2) steady state over K back-to-back steps: the H2D of step i+1 and the D2H of step i-1 overlap the kernels
        # of step i (polycracker_b200.overlap.StepPipeline: three streams, double-buffered staging).  Every step's
        # genome still crosses PCIe from pinned host memory and every step's CSR + TF-IDF still lands in pinned host
        # memory inside the timed region.
        from polycracker_b200.overlap import StepPipeline
        pipe = StepPipeline(ctx, device)
        host_in = pinned[:local_bases]

        def run_dev(dev_src):
            if world == 1:
                return run_hot_path(ctx, dev_src, offs_all, names_all, chunk_size=chunk, k=k, kmer_low_count=low,
                                    fetch=False, layout=layout)
            return pcd.run_hot_path_distributed(ctx, layout, None, device, chunk_size=chunk, k=k, kmer_low_count=low,
                                                fetch=False, staged=(dev_src, ids, begins, ends),
                                                exchange=os.environ.get("PC_EXCHANGE", "auto"))

        e2e_wall = []

        def piped(n_steps):
            del e2e_wall[:]
            gc.disable()
            try:
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pipe.prefetch(host_in)
                out = None
                for i in range(n_steps):
                    t_s = time.perf_counter()
                    if i + 1 < n_steps:
                        pipe.prefetch(host_in)
                    got = pipe.step(run_dev)
                    out = got if got is not None else out
                    e2e_wall.append((time.perf_counter() - t_s) * 1e3)
                last = pipe.flush()
                e1.record()
                barrier()
                ms = e0.elapsed_time(e1)
                if dist is not None:
                    t = torch.tensor([ms], dtype=torch.float64, device=device)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                return ms, last
            finally:
                gc.enable()

        piped(2)
        h0, d0 = pipe.h2d_bytes, pipe.d2h_bytes
        e2e_attempts = []
        for attempt in range(3):                           # same re-measure rule as the resident loop (3 x the median here:
            ms_e2e, res_p = piped(args.steps)              # the first and last steps of a pipelined loop are longer by design)
            stalled = len(e2e_wall) >= 3 and max(e2e_wall) > 3.0 * float(np.median(e2e_wall))
            if dist is not None:
                f = torch.tensor([1.0 if stalled else 0.0], dtype=torch.float64, device=device)
                dist.all_reduce(f, op=dist.ReduceOp.MAX)
                stalled = bool(f.item() > 0)
            e2e_attempts.append((ms_e2e, res_p, list(e2e_wall), stalled))
            if not stalled:
                break
        clean = [a for a in e2e_attempts if not a[3]]
        ms_e2e, res_p, e2e_wall, _ = clean[0] if clean else min(e2e_attempts, key=lambda a: a[0])
        pipe.close()
        h2d = (pipe.h2d_bytes - h0) // (args.steps * len(e2e_attempts))
        d2h = (pipe.d2h_bytes - d0) // (args.steps * len(e2e_attempts))
        same = (np.array_equal(res_p.indptr, res.indptr) and np.array_equal(res_p.indices, res.indices)
                and np.array_equal(res_p.data, res.data) and np.array_equal(res_p.kmer_keys, res.kmer_keys)
                and np.array_equal(res_p.tfidf, res.tfidf))
        if not same:
            raise SystemExit("pipelined end-to-end result differs from the sequential one")
        e2e = {"value": genome_bases / (ms_e2e / args.steps * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
               "sequential_ms_per_step": ms_lat / args.steps,
               "sequential_value": genome_bases / (ms_lat / args.steps * 1e-3) / 1e9,
               "timed_attempts": [{"ms_per_step": round(a[0] / args.steps, 3), "stalled_step": a[3]} for a in e2e_attempts],
               "note": "K back-to-back steps through the public API, host buffers: pinned host genome -> H2D -> "
                       "pc_load_sequences ... pc_csr_get8/pc_tfidf_factors_get -> D2H into pinned host memory, every step "
                       "(compact hand-off: one byte per column gap + one per count + escape lists + idf + row norms, ~2 B per entry; the int32 / float32 "
                       "count / tf-idf arrays are rebuilt on the host bit for bit, outside the timed region); copies "
                       "of neighbouring steps overlap this step's kernels (StepPipeline).  sequential_* = the same "
                       "with nothing overlapped (latency of one call)" + ("; bytes per rank" if world > 1 else "")}

    numa_all = [numa]
    if dist is not None:
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)
    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel, from the CUDA-event time of its launches ------------------------------
    # ALGORITHMIC bytes (DESIGN.md "Roofline model"), per rank: G packed bases, T valid windows,
    # cap table slots, Ds selected k-mers, nnz matrix entries, R rows.  The K3 model (0.375 G + 24 T) is split over
    # the three kernels that implement K3 here: the bucket count kernel owns the 24 T (one 12-byte slot
    # read-modify-write per k-mer instance), extract + scatter own the packed-genome read.
    peak, peak_src = hbm_peak()
    stage_ms = {name: float(np.mean([t[name] for t in per_step_timings])) for name in per_step_timings[0]}
    st = last_resident.stats
    G = sum(int((e - b) // 32 + 1) for b, e in zip(begins.tolist(), ends.tolist())) * 32
    T, cap, Ds, nnz, R = st["valid_windows"], info["capacity"], st["n_selected"], st["nnz"], st["n_rows"]
    detail = detail_main
    # K4 scans the count result: 12 B x table slots, or 12 B x list entries when the shared-memory count left a
    # compact (k-mer, count) list instead of a table
    scan_units = detail["kept"] if detail["kind"] == "list" else cap
    k3_parts = {"extract": stage_ms["part_hist"], "scatter_level1": stage_ms["part_scatter"], "scatter_level2": stage_ms["part2"],
                "table_init": stage_ms["table_init"], "expand": stage_ms.get("expand", 0.0), "count": stage_ms["count"],
                "count_fallback": stage_ms.get("count_fallback", 0.0)}
    k3_ms = sum(k3_parts.values())
    # one entry per unit of SURVEY.md 8d; K3 is ONE unit (0.375 G + 24 T) over all of its launches
    model = {
        "K1_pack": (1.375 * G, stage_ms["pack"]),
        "K3_count": (0.375 * G + 24.0 * T, k3_ms),
        "K4_select+sort": (12.0 * scan_units + 12.0 * Ds, stage_ms["select"] + stage_ms["sort"] + stage_ms["rep_table"]),
        "K5_probe": (0.375 * G + 12.0 * T + 8.0 * nnz, stage_ms["probe"]),
        "K6_rows_to_csr": (16.0 * nnz + 8.0 * R, stage_ms["row_sort"] + stage_ms["emit"]),
        "K7_tfidf": (12.0 * nnz + 4.0 * Ds, stage_ms["tfidf"]),
    }
    kernels = {}
    for name, (nbytes, ms) in model.items():
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        kernels[name] = {"algorithmic_bytes": nbytes, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak}
    kernels["K3_count"]["launch_ms"] = {k_: round(v, 4) for k_, v in k3_parts.items()}
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    d = kernels[dom]
    path_bytes = sum(v[0] for v in model.values())
    skm = detail.get("skm_m", 0) > 0
    unit_kernels = {
        "K3_count": ("k_skm_extract + k_rec_scatter<1> + k_rec_scatter<2> + k_count_skm (super-k-mer count)" if skm else
                     "k_extract + k_tile_scatter + k_sub_scatter + k_count_smem (k-mer count)") if detail["kind"] == "list"
        else "k_extract + k_tile_scatter + k_count_part (table count)",
        "K5_probe": "k_probe_roll_filt" if skm else "k_probe_stream_filt", "K6_rows_to_csr": "k_row_bitmap_csr"}
    roofline = {"bound": "hbm", "kernel": unit_kernels.get(dom, dom),
                "achieved": d["achieved_gbs"],
                "peak": peak, "unit": "GB/s", "frac": d["frac"],
                "traffic": NCU_TRAFFIC.get((dom, detail["kind"], detail.get("skm_m", 0), int(genome_bases))),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"], "launch_ms": d["ms"],
                "whole_path": {"algorithmic_bytes": path_bytes, "ms": ms_per_step,
                               "achieved_gbs": path_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / peak},
                "per_kernel": kernels, "stage_ms": stage_ms,
                "count": detail,
                "model": "SURVEY.md 8d per-unit bytes x units of this step, per rank; the dominant unit is the one with the "
                         "largest CUDA-event time over ALL of its launches (K3 = extraction + both scatters + count + "
                         "fall-back, one unit with 0.375 G + 24 T).  `traffic` is null unless a committed ncu capture of this "
                         "workload and kernel variant exists (NCU_TRAFFIC)"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:            # the CPU baseline is reported at N = 1 only
        from oracle import c_oracle, ref_path
        sample_bases = args.cpu_sample or min(per_gpu, 250_000_000)     # ~10 s of CPU work on 16 cores
        r_cfg, r_params = ref_path.workload(WORKLOAD, args.genome_size)
        s_seq, s_off, s_names = ref_path.generate(r_params, max_bases=sample_bases, threads=8)
        dt, _ = cpu_reference_pass(s_seq, s_off, s_names, r_cfg)
        cpu_baseline = {"value": len(s_seq) / dt / 1e9, "unit": UNIT, "cores": c_oracle.threads(), "kind": "port",
                        "sample": "%d bp (first %d chromosomes of the %s genome), full path, %.1f s wall" % (
                            len(s_seq), len(s_names), WORKLOAD, dt)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_done": warmup_done,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "genome_bases": genome_bases, "bases_per_gpu": per_gpu, "k": k,
                       "chunk_size": chunk, "kmer_low_count": low, "n_subgenomes": cfg["n_subgenomes"],
                       "l2_policy": "inputs larger than L2 (ASCII %.0f MB, %s, count %s %.1f GB per rank)" % (
                           local_bases / 1e6,
                           "super-k-mer records %.1f GB" % (16.0 * detail.get("skm_records", 0) / 1e9) if skm else
                           "k-mer stream %.1f GB" % (8.0 * T / 1e9), "buffers" if detail["kind"] == "list" else "table",
                           info["table_bytes"] / 1e9),
                       "parallelism": ("one GPU" if world == 1 else
                                       "chunks sharded over %d ranks; k-mers hash-partitioned by owner, owners read the "
                                       "senders' bucket buffers over NVLink peer memory (level-2 histograms exchanged), "
                                       "all-gather of the repeat set, all-reduce of df" % world)},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "parity_check": parity,
            "gpu_launches": int(launches),
            "numa_binding": numa_all, "iso_work": iso, "clocks": clocks, "step_wall_ms": [round(x, 2) for x in step_wall],
            "timed_attempts": [{"ms_per_step": round(a["ms_per_step"], 3), "stalled_step": a["stalled_step"],
                                "max_step_wall_ms": round(max(a["walls"]), 2) if a["walls"] else None} for a in attempts],
            "e2e_step_wall_ms": [round(x, 2) for x in (e2e_wall if not args.no_e2e else [])], "quiet_gpu_wait_s": round(quiet_wait, 2),
            "result": {"distinct_kmers_rank0": info["distinct"], "valid_windows_rank0": info["valid_windows"],
                       "shared_tables": detail.get("sub_buckets"), "shared_table_slots": info["capacity"] // max(1, detail.get("sub_buckets") or 1),
                       "mean_shared_table_load": info["distinct"] / max(1, info["capacity"]),
                       "overflowed_sub_buckets": detail.get("overflowed"),
                       "n_selected": (res.stats["n_selected"] if res is not None else None),
                       "nnz_rank0": (res.stats["nnz"] if res is not None else None),
                       "rows_rank0": (res.stats["n_rows"] if res is not None else None),
                       "phase_ms_rank0": main_phase,
                       "exchange_bytes_rank0": last_resident.stats.get("exchange_bytes_sent"),
                       "exchange": last_resident.stats.get("exchange")}}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference_arm(a) if a.impl == "reference" else run_b200_arm(a))
