"""Experiment: do the preparation kernel of one batch and the render kernel of another overlap usefully?
Two generators on two streams render alternate batches; compare the aggregate with one generator on one stream."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ofdg_b200 as ofdg

B, W, H = 64, 512, 384
steps = 400


def make():
    g = ofdg.Generator(device=0, mode=7, max_batch=B)
    g.synth_textures(500, 2 * W, 2 * H, seed=3)
    ps = ofdg.ParamStream(7)
    prepared = [g.prepare(ps.generate(B)) for _ in range(4)]
    out = [torch.empty((B, c, H, W), device="cuda") for c in (3, 3, 2)]
    return g, prepared, out


def run(gens, streams):
    for w in range(20):
        for (g, p, o), s in zip(gens, streams):
            g.render_prepared(p[w % 4], *o, s.cuda_stream)
    torch.cuda.synchronize()
    t = time.time()
    for i in range(steps):
        for (g, p, o), s in zip(gens, streams):
            g.render_prepared(p[i % 4], *o, s.cuda_stream)
    torch.cuda.synchronize()
    dt = time.time() - t
    return len(gens) * steps * B / dt


a, b = make(), make()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("one generator, one stream: %.0f samples/s" % run([a], [s1]))
print("two generators, two streams: %.0f samples/s" % run([a, b], [s1, s2]))
hp = torch.cuda.Stream(priority=-1)
print("two generators, second stream high priority: %.0f samples/s" % run([a, b], [s1, hp]))
print("one generator, one stream: %.0f samples/s" % run([b], [s2]))
