"""Experiment: host-blob pipeline throughput against worker-thread and chunk count (run once per setting)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ofdg_b200 as ofdg

B, W, H = 64, 512, 384
g = ofdg.Generator(device=0, mode=7, max_batch=B)
g.synth_textures(1000, 2 * W, 2 * H, seed=3)
ps = ofdg.ParamStream(7)
h0 = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
h1 = torch.empty_like(h0).pin_memory()
hf = torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory()
for _ in range(5):
    g.generate_host(ps, B, h0, h1, hf)
best = 1e9
t_all = time.time()
N = 60
for _ in range(N):
    t = time.time()
    g.generate_host(ps, B, h0, h1, hf)
    best = min(best, time.time() - t)
avg = (time.time() - t_all) / N
print("threads=%s chunks=%s: avg %.2f ms (%.0f samples/s), best %.2f ms" % (
    os.environ.get("OFDG_HOST_THREADS", "default"), os.environ.get("OFDG_HOST_CHUNKS", "default"), avg * 1e3, B / avg, best * 1e3))
