"""Committed golden vectors (tests/golden/render_golden.npz, made by tests/golden/make_golden.py): the CPU oracle
must keep reproducing them (CPU suite), and the sm_100a path must reproduce them through the C ABI (GPU suite)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = np.load(os.path.join(HERE, "golden", "render_golden.npz"))


def _check(mode, got, flow_tol):
    g = lambda k: GOLD[f"m{mode}_{k}"]
    assert got["frames_sha"] == str(g("frames_sha")), "uint8 frames differ from the golden vectors"
    assert got["id0_sha"] == str(g("id0_sha")) and got["id1_sha"] == str(g("id1_sha")), "index images differ"
    assert np.array_equal(got["probes"], g("probes"))
    assert np.abs(got["flow"] - g("flow")).max() <= flow_tol
    assert np.abs(got["flow_bw"] - g("flow_bw")).max() <= flow_tol


@pytest.mark.parametrize("mode,n", [(1, 2), (7, 2)])
def test_oracle_reproduces_golden_vectors(ofdg, oracle, mode, n):
    _check(mode, mg.compute(mode, n), 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,n", mg.CASES)
def test_gpu_reproduces_golden_vectors(ofdg, mode, n):
    """Frames and index images bit-exact (1 LSB is allowed by the spec, 0 is what the kernels deliver), flow within 1e-3 px."""
    import torch
    g = ofdg.Generator(device=0, mode=mode, max_batch=n)
    g.upload_textures(ofdg.synth_textures(8, 1024, 768, seed=1))
    tasks = ofdg.ParamStream(mode).generate(n)
    dbg = g.render_debug(tasks, want_masks=False)
    bw = torch.empty((n, 2, 384, 512), device="cuda")
    g.set_extra_tops(flow_bw=bw)
    i0 = torch.empty((n, 3, 384, 512), device="cuda"); i1 = torch.empty_like(i0); fl = torch.empty((n, 2, 384, 512), device="cuda")
    g.render(tasks, i0, i1, fl)
    torch.cuda.synchronize()
    out = {"frames8": dbg["frames8"], "id0": dbg["id0"], "id1": dbg["id1"], "flow": fl.cpu().numpy(), "flow_bw": bw.cpu().numpy()}
    _check(mode, mg.summarise(out), 1e-3)
    g.close()
