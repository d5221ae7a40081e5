"""Dumps the device parameter stream's blueprints (and one rendered batch) to an .npz, to compare two builds of the library:
python tools/dump_philox.py out.npz [compare_with.npz]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ofdg_b200 as o

out = {}
for mode in (1, 3, 5, 7, 9, 13):
    g = o.Generator(device=0, mode=mode, max_batch=32)
    g.synth_textures(8, 1024, 768, seed=2)
    for aug in (False, True):
        arrs = g.philox_tasks(1234, 100, 32, augment=aug).arrays()
        for k, v in arrs.items():
            if v is not None:
                out["m%d_a%d_%s" % (mode, aug, k)] = np.frombuffer(np.ascontiguousarray(v).tobytes(), dtype=np.uint8)
    if mode == 7:
        a = torch.empty(32, 3, 384, 512, device="cuda"); b = torch.empty_like(a); f = torch.empty(32, 2, 384, 512, device="cuda")
        g.generate_philox(1234, 100, 32, a, b, f)
        torch.cuda.synchronize()
        out["m7_img1_sum"] = np.array([float(b.double().sum()), float(f.double().abs().sum())])
np.savez(sys.argv[1], **out)
if len(sys.argv) > 2:
    ref = np.load(sys.argv[2])
    bad = [k for k in out if k not in ref or not np.array_equal(ref[k], out[k])]
    print("philox stream identical to", sys.argv[2], ":", not bad, bad[:5])
