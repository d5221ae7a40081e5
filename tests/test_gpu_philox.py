"""Device-side (Philox) parameter stream + device flattening: production mode.
Statistically equivalent to the host stream, bitwise reproducible from (seed, sample index)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gen(ofdg, mode, **kw):
    import torch
    assert torch.cuda.is_available()
    return ofdg.Generator(device=0, mode=mode, max_batch=64, **kw)


def _blobs(n):
    import torch
    return (torch.empty((n, 3, 384, 512), device="cuda"), torch.empty((n, 3, 384, 512), device="cuda"),
            torch.empty((n, 2, 384, 512), device="cuda"))


@pytest.mark.parametrize("mode", [2, 5, 7])
def test_device_stream_renders_what_it_draws(ofdg, oracle, textures8, mode):
    g = _gen(ofdg, mode)
    g.upload_textures(textures8)
    i0, i1, fl = _blobs(8)
    g.generate_philox(1234, 100, 8, i0, i1, fl)
    dev = [t.cpu().numpy() for t in (i0, i1, fl)]
    # the same scenes as an ordinary task batch: the host path and the oracle agree on them exactly ...
    tasks = g.philox_tasks(1234, 100, 8)
    host = g.render_host(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=mode)
    assert np.abs(host[0] - cpu["img0"]).max() <= 1 and np.abs(host[1] - cpu["img1"]).max() <= 1
    assert np.abs(host[2] - cpu["flow"]).max() <= 1e-3
    # ... and the device-flattened render differs from them only where the device libm rounds a vertex differently
    for d, h in zip(dev, host):
        bad = np.abs(d - h) > (1e-3 if d.shape[1] == 2 else 1)
        assert bad.mean() < 2e-3, bad.mean()
    assert dev[0].std() > 10
    g.close()


def test_sample_is_a_pure_function_of_seed_and_index(ofdg, textures8):
    import torch
    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    a = _blobs(8)
    b = _blobs(4)
    g.generate_philox(7, 0, 8, *a)
    g.generate_philox(7, 4, 4, *b)        # another "rank" rendering samples 4..7
    for x, y in zip(a, b):
        assert torch.equal(x[4:8], y)
    c = _blobs(4)
    g.generate_philox(8, 4, 4, *c)        # another seed: other samples
    assert not torch.equal(b[0], c[0])
    g.close()


def test_batches_drawn_ahead_are_used_or_discarded_correctly(ofdg, textures8):
    """The generator draws the next two batches of a stream ahead (three scene sets). Consecutive calls consume them in order;
    a call that asks for anything else (another position, another seed, another batch size) discards them. Either way the
    blobs equal those of a generator that was asked for that batch first thing."""
    import torch

    def fresh(seed, first, n):
        g = _gen(ofdg, 7)
        g.upload_textures(textures8)
        out = _blobs(n)
        g.generate_philox(seed, first, n, *out)
        torch.cuda.synchronize()
        g.close()
        return out

    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    got = {}
    for call, (seed, first, n) in enumerate([(5, 0, 4), (5, 4, 4), (5, 8, 4), (5, 12, 4),     # in order: the drawn-ahead batches are used
                                             (5, 4, 4),                                         # back: discarded
                                             (5, 8, 4),                                         # in order again
                                             (6, 8, 4), (6, 12, 4),                             # another seed
                                             (6, 16, 2), (6, 18, 2), (6, 20, 2)]):              # another batch size
        out = _blobs(n)
        g.generate_philox(seed, first, n, *out)
        torch.cuda.synchronize()
        got[call] = (seed, first, n, [t.clone() for t in out])
    g.close()
    want = {}
    for call, (seed, first, n, out) in got.items():
        key = (seed, first, n)
        if key not in want:
            want[key] = fresh(*key)
        for x, y in zip(out, want[key]):
            assert torch.equal(x, y), (call, key)


def test_statistics_match_the_host_stream(ofdg, textures8):
    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    dev = g.philox_tasks(99, 0, 64).arrays()["blueprints"]
    for k in range(1, 4):
        dev = np.concatenate([dev, g.philox_tasks(99, 64 * k, 64).arrays()["blueprints"]])
    host = ofdg.ParamStream(7).generate(256).arrays()["blueprints"]

    def stats(bp):
        top = bp[(bp["parent"] < 0) & (bp["obj_id"] >= 10)]
        bg = bp[bp["obj_id"] == 1]
        comps = bp[bp["parent"] >= 0]
        return {
            "fg_per_sample": len(top) / len(bg),
            "ellipse": (top["obj_type"] == 1).mean(), "polygon": (top["obj_type"] == 2).mean(), "composite": (top["obj_type"] == 3).mean(),
            "rot_on": (top["rot"] != 0).mean(), "scale_on": (top["scale"] != 1).mean(),
            "bg_rot_on": (bg["rot"] != 0).mean(), "trans_std": top["trans_x"].std(), "init_x_mean": top["init_trans_x"].mean(),
            "comps_per_composite": len(comps) / max(1, (top["obj_type"] == 3).sum()),
            "ellipse_scale_y": top[top["obj_type"] == 1]["ellipse_scale_y"].mean(),
        }
    sd, sh = stats(dev), stats(host)
    tol = {"fg_per_sample": 0.5, "ellipse": 0.04, "polygon": 0.04, "composite": 0.04, "rot_on": 0.04, "scale_on": 0.04,
           "bg_rot_on": 0.1, "trans_std": 4.0, "init_x_mean": 20.0, "comps_per_composite": 0.4, "ellipse_scale_y": 3.0}
    for k in sd:
        assert abs(sd[k] - sh[k]) <= tol[k], (k, sd[k], sh[k])
    top = dev[(dev["parent"] < 0) & (dev["obj_id"] >= 10)]
    assert np.abs(top["trans_x"]).max() <= 120 and np.abs(top["rot"]).max() <= 30 * np.pi / 180 + 1e-6
    assert top["init_trans_x"].min() >= -306 and top["init_trans_x"].max() <= 818
    g.close()


def test_device_stream_with_augmentation(ofdg, textures8):
    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    a, b = _blobs(4), _blobs(4)
    g.generate_philox(5, 0, 4, *a)
    g.generate_philox(5, 0, 4, *b, augment=True)
    assert (a[0] - b[0]).abs().mean().item() > 2 and (a[2] - b[2]).abs().max().item() == 0
    g.close()


def test_device_stream_mode9_warp_fields(ofdg, oracle, textures8):
    """Mode 9 through the device stream: field picks are counter-based draws, the flatten kernel hands out the deformation
    slots, and the render equals the host path's render of the very same blueprints."""
    import torch
    fields = oracle.generate_fields(512, 384, seed=3, n_fields=4)
    g = _gen(ofdg, 9)
    g.upload_textures(textures8)
    g.set_fields(fields)
    n = 12
    tasks = g.philox_tasks(77, 40, n)
    bp = tasks.arrays()["blueprints"]
    flagged = bp[(bp["do_warpfield_deformation"] != 0) & (bp["parent"] < 0)]
    assert len(flagged) >= 3 and set(np.unique(flagged["field_id"])) <= {0, 1, 2, 3} and len(np.unique(flagged["field_id"])) >= 2
    assert (flagged["obj_id"] == 1).any(), "a deformed background must be part of the batch"
    host = g.render_host(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=9, fields=fields)
    ok = np.isfinite(cpu["flow"])
    assert np.abs(host[0] - cpu["img0"]).max() <= 1 and np.abs(host[2][ok] - cpu["flow"][ok]).max() <= 1e-3
    for rep in range(2):          # the second call is served by the look-ahead set
        i0, i1, fl = _blobs(n)
        g.generate_philox(77, 40 + rep * n, n, i0, i1, fl)
        torch.cuda.synchronize()
        if rep == 0:
            dev = [t.cpu().numpy() for t in (i0, i1, fl)]
            for d, h in zip(dev, host):
                with np.errstate(invalid="ignore"):
                    bad = np.abs(d - h) > (1e-3 if d.shape[1] == 2 else 1)
                assert bad.mean() < 3e-3, bad.mean()
        else:
            nxt = g.render_host(g.philox_tasks(77, 40 + n, n))
            with np.errstate(invalid="ignore"):
                bad = np.abs(i0.cpu().numpy() - nxt[0]) > 1
            assert bad.mean() < 3e-3
    g.close()


def test_device_stream_reports_truncated_scenes(ofdg, textures8, monkeypatch):
    """The device stream works in fixed strides (32 objects per sample, 512 vertices per outline and frame). A scene beyond them
    used to lose objects silently; now the call that notices fails loudly (OFDG_TEST_PHILOX_FG forces 40 objects)."""
    import torch
    monkeypatch.setenv("OFDG_TEST_PHILOX_FG", "40")
    g = ofdg.Generator(device=0, mode=7, max_batch=4)
    g.upload_textures(textures8)
    i0 = torch.empty((4, 3, 384, 512), device="cuda")
    i1 = torch.empty_like(i0)
    fl = torch.empty((4, 2, 384, 512), device="cuda")
    with pytest.raises(ofdg.OfdgError, match="more than 32 foreground objects"):
        g.generate_philox(1, 0, 4, i0, i1, fl)
    monkeypatch.delenv("OFDG_TEST_PHILOX_FG")
    g.close()
    g = ofdg.Generator(device=0, mode=7, max_batch=4)
    g.upload_textures(textures8)
    g.generate_philox(1, 0, 4, i0, i1, fl)  # the ordinary stream: fine
    assert float(i0.std()) > 5
    g.close()


@pytest.mark.gpu
def test_device_stream_is_pinned(ofdg):
    """The blueprints of fixed (mode, seed, first sample) cases hash to the committed digests (tests/golden/philox_digest.json,
    made by tests/golden/make_philox_digest.py): rewriting the generating kernels (lanes over a polygon's spokes, staging in
    shared memory) must not move the stream."""
    import importlib.util
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_philox_digest", os.path.join(here, "golden", "make_philox_digest.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    want = json.load(open(os.path.join(here, "golden", "philox_digest.json")))
    got = m.digests()
    assert got == want
