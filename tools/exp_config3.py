"""Config 3 on its own (mode 9 + generated warp fields + augmentation, batch 64), for launch lists / ncu captures:
    python tools/exp_config3.py [steps] [augment 0|1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ofdg_b200 as o
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
aug = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = o.Generator(device=0, width=512, height=384, mode=9, max_batch=64)
g.synth_textures(1000, 1024, 768, seed=0)
g.generate_fields(5, 40)
ps = o.ParamStream(9, 512, 384, n_fields=40)
ps.enable_augmentation(bool(aug))
prepared = [g.prepare(ps.generate(64)) for _ in range(3)]
img0 = torch.empty((64, 3, 384, 512), device="cuda"); img1 = torch.empty_like(img0); flow = torch.empty((64, 2, 384, 512), device="cuda")
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for i in range(4):
    g.render_prepared(prepared[i % 3], img0, img1, flow, ts.cuda_stream)
torch.cuda.synchronize()
g.kernel_times()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    g.render_prepared(prepared[i % 3], img0, img1, flow, ts.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
prep, render, calls = g.kernel_times()
print("config3 augment=%d: %.3f ms/step, %.0f samples/s; bg_prep %.3f render %.3f shade %.3f (per call)" % (aug, ms, 64 / ms * 1e3, prep / calls, render / calls, g.last_shade_ms() / calls))
