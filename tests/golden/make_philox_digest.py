"""Digests of the device-side parameter stream (ofdg_philox_tasks) for fixed (mode, seed, first sample): run on a B200,
commit the output as tests/golden/philox_digest.json. The stream is a pure function of (mode, seed, sample index) -- the
digests pin it against changes of the generating kernels (tests/test_gpu_philox.py::test_device_stream_is_pinned).
    python tests/golden/make_philox_digest.py out.json"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import ofdg_b200 as o

CASES = [(1, 11, 0), (3, 12, 64), (5, 13, 128), (7, 1234, 100), (9, 77, 5), (13, 99, 1000)]


def digests():
    out = {}
    for mode, seed, first in CASES:
        g = o.Generator(device=0, mode=mode, max_batch=32)
        g.synth_textures(8, 1024, 768, seed=2)
        if mode == 9:
            g.generate_fields(3, 4)
        for aug in (0, 1):
            arrs = g.philox_tasks(seed, first, 32, augment=bool(aug)).arrays()
            h = hashlib.sha256()
            for k in sorted(arrs):
                if arrs[k] is not None:
                    h.update(k.encode())
                    h.update(np.ascontiguousarray(arrs[k]).tobytes())
            out["mode%d_seed%d_first%d_aug%d" % (mode, seed, first, aug)] = h.hexdigest()
    return out


if __name__ == "__main__":
    json.dump(digests(), open(sys.argv[1], "w"), indent=1, sort_keys=True)
