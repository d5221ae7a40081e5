"""Experiment: stage timestamps of the host-blob pipeline (OFDG_TRACE_HOST=1 python tools/exp_host_trace.py)."""
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
for _ in range(12):
    t = time.time()
    g.generate_host(ps, B, h0, h1, hf)
    print("call %.2f ms" % ((time.time() - t) * 1e3), file=sys.stderr)
tasks = ofdg.ParamStream(7).generate(B)
for _ in range(4):
    t = time.time()
    g.render_host(tasks, h0, h1, hf)
    print("render_host (tasks given) %.2f ms" % ((time.time() - t) * 1e3), file=sys.stderr)
