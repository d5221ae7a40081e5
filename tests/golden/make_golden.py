"""Generates tests/golden/render_golden.npz for fixed parameter streams and procedural textures. The reference ships no
fixtures of its own. The vectors are computed by the CPU oracle (oracle/, the restatement of the reference's
render path); tests/test_reference_build.py::test_golden_vectors_are_reference_outputs shows that the reference's own code
(oracle/_ref: its untouched sources compiled against stand-in AGG / CImg headers) renders the very same bytes, so the file
pins reference outputs -- up to the third-party arithmetic both sides restate. Any later change to the oracle, to the host
parameter stream or to the texture synthesis that alters a pixel shows up as a diff against this file, and the sm_100a path
is checked against the same vectors on the GPU box. `--ref` computes them with the reference build instead.

    python tests/golden/make_golden.py        # rewrites render_golden.npz

Stored per case (2 samples at 512 x 384): SHA-256 of the uint8 frames and of the two index images, the forward and
backward flow sampled every 8th pixel, and 64 probe pixels of the frames.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [(1, 2), (2, 2), (5, 2), (7, 2), (12, 2)]  # (mode, samples)
PROBES = np.random.default_rng(7).integers(0, 384 * 512, 64)


def compute(mode, n, use_ref=False):
    import ofdg_b200 as ofdg
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    tasks = ofdg.ParamStream(mode).generate(n)
    if use_ref:
        from oracle import ref_binding
        return summarise(ref_binding.Generator(mode, textures=tex).render(tasks.struct(), debug=True))
    from oracle import binding as oracle
    oracle.build()
    out = oracle.render(tasks.struct(), tex, mode=mode, debug=True)
    return summarise(out)


def summarise(out):
    f8 = np.ascontiguousarray(out["frames8"])
    return {
        "frames_sha": hashlib.sha256(f8.tobytes()).hexdigest(),
        "id0_sha": hashlib.sha256(np.ascontiguousarray(out["id0"], np.uint32).tobytes()).hexdigest(),
        "id1_sha": hashlib.sha256(np.ascontiguousarray(out["id1"], np.uint32).tobytes()).hexdigest(),
        "flow": np.ascontiguousarray(out["flow"][:, :, ::8, ::8]),
        "flow_bw": np.ascontiguousarray(out["flow_bw"][:, :, ::8, ::8]),
        "probes": f8.reshape(f8.shape[0], 6, -1)[:, :, PROBES].copy(),
    }


if __name__ == "__main__":
    blob = {}
    for mode, n in CASES:
        s = compute(mode, n, use_ref="--ref" in sys.argv)
        for k, v in s.items():
            blob[f"m{mode}_{k}"] = np.array(v) if isinstance(v, str) else v
        print("mode", mode, s["frames_sha"][:16], s["id0_sha"][:16])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "render_golden.npz"), **blob)
