"""Development probe: how much do NVML queries (clock / throttle reasons) stall CUDA work on this box?"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
from polycracker_b200 import binding, synth
from polycracker_b200.pipeline import ChunkLayout, run_hot_path

pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
p = synth.make_params(20260002, 200_000_000); seq, off, names = synth.generate(p, threads=8)
lay = ChunkLayout(names, off, 75000); lay.rows()
dev = torch.from_numpy(seq).cuda()
ctx = binding.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
def step(): run_hot_path(ctx, dev, off, names, chunk_size=75000, k=23, kmer_low_count=30, fetch=False, layout=lay)
for _ in range(3): step()
get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
modes = {"none": None, "clock": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
         "reasons": lambda: get_reasons(h), "power": lambda: pynvml.nvmlDeviceGetPowerUsage(h)}
for name, fn in modes.items():
    stop = False; durs = []
    def loop():
        while not stop:
            t = time.perf_counter(); fn(); durs.append((time.perf_counter() - t) * 1e3); time.sleep(0.05)
    th = None
    if fn: th = threading.Thread(target=loop, daemon=True); th.start()
    walls = []
    for _ in range(12):
        t = time.perf_counter(); step(); torch.cuda.synchronize(); walls.append((time.perf_counter() - t) * 1e3)
    stop = True
    if th: th.join()
    print(name, "step ms", [round(w, 1) for w in walls], "query ms", [round(d, 2) for d in durs[:8]], flush=True)
