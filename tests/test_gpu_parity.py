"""GPU parity: the sm_100a render path (through the C ABI) against the CPU oracle on identical
task batches and textures. Thresholds (BASELINE.json north_star / SURVEY 8d):
  coverage masks and object-id images bit-exact; uint8 images within 1 LSB; flow within 1e-3 px."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOW_TOL = 1e-3
IMG_TOL = 1  # LSB


def _gen(ofdg, mode, n_tex=8, max_batch=16, **kw):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return ofdg.Generator(device=0, mode=mode, max_batch=max_batch, **kw)


def _compare(gpu, cpu, n_objs_max=24):
    assert np.array_equal(gpu["id0"], cpu["id0"]), "object-id image (frame 0) differs"
    assert np.array_equal(gpu["id1"], cpu["id1"]), "object-id image (frame 1) differs"
    assert np.array_equal(gpu["masks"], cpu["masks"]), "coverage masks differ"
    d = np.abs(gpu["frames8"].astype(np.int32) - cpu["frames8"].astype(np.int32)).max()
    assert d <= IMG_TOL, f"uint8 frames differ by {d} LSB"
    assert np.abs(gpu["img0"] - cpu["img0"]).max() <= IMG_TOL
    assert np.abs(gpu["img1"] - cpu["img1"]).max() <= IMG_TOL
    assert np.abs(gpu["flow"] - cpu["flow"]).max() <= FLOW_TOL
    return int(d)


@pytest.mark.parametrize("mode", [1, 2, 3, 5, 7])
def test_mode_parity_batch8(ofdg, oracle, textures8, mode):
    g = _gen(ofdg, mode)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(mode).generate(8)
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=mode, debug=True)
    d = _compare(gpu, cpu)
    # the float blobs are the uint8 frames converted, channel-major (NCHW)
    assert np.array_equal(gpu["img0"], gpu["frames8"][:, 0].astype(np.float32))
    assert np.array_equal(gpu["img1"], gpu["frames8"][:, 1].astype(np.float32))
    print(f"mode {mode}: max image diff {d} LSB")
    g.close()


def test_no_antialiasing(ofdg, oracle, textures8):
    g = _gen(ofdg, 7, use_antialiasing=False)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(7).generate(4)
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=7, use_aa=False, debug=True)
    _compare(gpu, cpu)
    g.close()


def test_background_preparation(ofdg, oracle, textures8):
    """Texture::getRandomizedCrop(2W, 2H, rot, zoom, shift) for the background, inside the region the renderer reads."""
    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(7).generate(8)
    a = tasks.arrays()
    bg, need = g.debug_background(tasks)
    zooms = []
    for t in range(8):
        b = a["blueprints"][a["task_begin"][t]]
        ref = oracle.randomized_crop(textures8[b["tex_id"] % 8], 1024, 768, float(b["tex_rot"]), float(b["tex_scale"]),
                                     int(b["tex_shift_x"]), int(b["tex_shift_y"]))
        x0, y0, x1, y1 = need[t]
        assert np.array_equal(bg[t][:, y0:y1 + 1, x0:x1 + 1], ref[:, y0:y1 + 1, x0:x1 + 1]), f"task {t} zoom {b['tex_scale']}"
        zooms.append(float(b["tex_scale"]))
    assert min(zooms) < 1 < max(zooms), "both the shrinking (moving average) and growing (linear) resize paths must be hit"
    g.close()


def test_synth_textures_match_numpy(ofdg):
    g = _gen(ofdg, 1)
    g.synth_textures(3, 1024, 768, seed=5)
    ref = ofdg.synth_textures(3, 1024, 768, seed=5)
    for i in range(3):
        assert np.array_equal(g.download_texture(i), ref[i])
    g.close()


def test_upload_roundtrip(ofdg, textures8):
    g = _gen(ofdg, 1)
    g.upload_textures(textures8)
    assert np.array_equal(g.download_texture(5), textures8[5])
    g.close()


def test_device_blobs_and_prepared(ofdg, oracle, textures8):
    """ofdg_render into torch device tensors == ofdg_render_host == ofdg_render_prepared."""
    import torch
    g = _gen(ofdg, 5)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(5).generate(8)
    i0 = torch.empty((8, 3, 384, 512), device="cuda"); i1 = torch.empty_like(i0)
    fl = torch.empty((8, 2, 384, 512), device="cuda")
    g.render(tasks, i0, i1, fl)
    h0, h1, hf = g.render_host(tasks)
    assert np.array_equal(i0.cpu().numpy(), h0) and np.array_equal(i1.cpu().numpy(), h1) and np.array_equal(fl.cpu().numpy(), hf)
    p = g.prepare(tasks)
    j0 = torch.zeros_like(i0); j1 = torch.zeros_like(i1); jf = torch.zeros_like(fl)
    g.render_prepared(p, j0, j1, jf)
    torch.cuda.synchronize()
    assert torch.equal(i0, j0) and torch.equal(i1, j1) and torch.equal(fl, jf)
    cpu = oracle.render(tasks.struct(), textures8, mode=5)
    assert np.abs(h0 - cpu["img0"]).max() <= IMG_TOL and np.abs(hf - cpu["flow"]).max() <= FLOW_TOL
    assert g.launch_count() > 0
    g.close()


def test_composite_rules_exhaustive(ofdg, oracle):
    """All 2 x 65536 (u, v) pairs of the composite-mask float rules, device vs the reference's expression."""
    g = _gen(ofdg, 7)
    ga, gs = g.debug_composite_luts()
    ca, cs = oracle.composite_luts()
    assert np.array_equal(ga, ca) and np.array_equal(gs, cs)
    g.close()


def test_many_objects_per_tile(ofdg, oracle, textures8):
    """More objects than one pass of the tile kernel holds (stress config: forced object count)."""
    g = _gen(ofdg, 7, max_batch=2)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(7, fg_override=160).generate(2)
    gpu = g.render_debug(tasks, max_objs=160)
    cpu = oracle.render(tasks.struct(), textures8, mode=7, debug=True, max_objs=160)
    _compare(gpu, cpu)
    g.close()


@pytest.fixture(scope="module")
def fields4(oracle):
    f = oracle.generate_fields(512, 384, seed=1, n_fields=4)
    # the reference's fields carry NaNs near the canvas border (WarpFields.cpp:389-398): plant some
    f[3, :, :, 100:140, 200:260] = np.nan
    return f


def test_mode9_nonrigid_parity(ofdg, oracle, textures8, fields4):
    """Mode 9: warped masks / textures / flow for flagged objects, composites and backgrounds."""
    g = _gen(ofdg, 9)
    g.upload_textures(textures8)
    g.set_fields(fields4)
    tasks = ofdg.ParamStream(9, n_fields=4).generate(8)
    bp = tasks.arrays()["blueprints"]
    flagged = bp[(bp["do_warpfield_deformation"] != 0) & (bp["parent"] < 0)]
    assert (flagged["obj_id"] == 1).any(), "a deformed background must be part of the batch"
    assert (flagged["obj_type"] == 3).any() and (flagged["obj_type"] == 1).any() and (flagged["obj_type"] == 2).any()
    assert (flagged["field_id"] == 3).any(), "the field with NaNs must be in use"
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=9, fields=fields4, debug=True)
    assert np.array_equal(gpu["id0"], cpu["id0"]) and np.array_equal(gpu["id1"], cpu["id1"])
    assert np.array_equal(gpu["masks"], cpu["masks"])
    assert np.abs(gpu["frames8"].astype(int) - cpu["frames8"].astype(int)).max() <= IMG_TOL
    ok = np.isfinite(cpu["flow"])
    assert np.array_equal(ok, np.isfinite(gpu["flow"]))  # NaN flow where the field is NaN, exactly like the reference
    assert np.abs(gpu["flow"][ok] - cpu["flow"][ok]).max() <= FLOW_TOL
    g.close()


def test_double_resolution(ofdg, oracle):
    """Stress config: 2x output resolution (runtime W, H; the reference only has #defines)."""
    W, H = 1024, 768
    tex = ofdg.synth_textures(3, 2 * W, 2 * H, seed=11)
    g = ofdg.Generator(device=0, width=W, height=H, mode=7, max_batch=2)
    g.upload_textures(tex)
    tasks = ofdg.ParamStream(7, W, H, fg_override=40).generate(2)
    gpu = g.render_debug(tasks, max_objs=40)
    cpu = oracle.render(tasks.struct(), tex, W=W, H=H, mode=7, debug=True, max_objs=40)
    _compare(gpu, cpu)
    g.close()


def test_augmentation_bit_exact(ofdg, oracle, textures8):
    """Fused colour/noise augmentation (this repository's spec, not the reference's): integer Philox and
    single-rounded float operations on both sides, so the float blobs must agree exactly."""
    g = _gen(ofdg, 7)
    g.upload_textures(textures8)
    ps = ofdg.ParamStream(7)
    ps.enable_augmentation(True)
    tasks = ps.generate(8)
    aug = tasks.arrays()["augment"]
    assert aug is not None and np.all(aug["enabled"] == 1) and aug["noise_sigma"].max() > 1
    i0, i1, fl = g.render_host(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=7)
    assert np.array_equal(i0, cpu["img0"]) and np.array_equal(i1, cpu["img1"])
    assert np.abs(fl - cpu["flow"]).max() <= FLOW_TOL
    plain = oracle.render(ofdg.ParamStream(7).generate(8).struct(), textures8, mode=7)
    assert np.array_equal(plain["flow"], cpu["flow"])          # geometry untouched
    assert np.abs(plain["img0"] - cpu["img0"]).mean() > 3      # colours are not
    assert i0.min() >= 0 and i0.max() <= 255
    g.close()


def test_gpu_field_producer_matches_the_cpu_restatement(ofdg, oracle, textures8):
    """WarpFields::CropGenerator on the GPU vs oracle/warpfields.cpp, same seed (differences: device expf only)."""
    g = _gen(ofdg, 9)
    g.upload_textures(textures8)
    gpu = g.generate_fields(seed=1, n=6)
    cpu = oracle.generate_fields(512, 384, seed=1, n_fields=6)
    assert gpu.shape == cpu.shape == (6, 2, 2, 385, 513)
    assert np.array_equal(np.isnan(gpu), np.isnan(cpu))
    ok = np.isfinite(cpu)
    assert np.abs(gpu[ok] - cpu[ok]).max() < 2e-3 and np.abs(cpu[ok]).max() > 10
    # forward and inverse fields undo each other: x + flow(x) + iflow(x + flow(x)) ~= x
    f, fi = gpu[0, 0], gpu[0, 1]
    ys, xs = np.mgrid[100:300:7, 100:400:7]
    tx, ty = xs + f[0, ys, xs], ys + f[1, ys, xs]
    ix, iy = np.clip(np.round(tx).astype(int), 0, 512), np.clip(np.round(ty).astype(int), 0, 384)
    assert np.abs(tx + fi[0, iy, ix] - xs).max() < 1.5 and np.abs(ty + fi[1, iy, ix] - ys).max() < 1.5
    # the installed pool renders (parity of the consumer side is covered by test_mode9_nonrigid_parity)
    tasks = ofdg.ParamStream(9, n_fields=6).generate(2)
    out = g.render_debug(tasks)
    ref = oracle.render(tasks.struct(), textures8, mode=9, fields=gpu, debug=True)
    assert np.array_equal(out["masks"], ref["masks"]) and np.abs(out["frames8"].astype(int) - ref["frames8"].astype(int)).max() <= 1
    g.close()


@pytest.mark.parametrize("n", [1, 5, 17, 64])
def test_host_blobs_uint8_transport_equals_float_transport(ofdg, textures8, n, monkeypatch):
    """The host-blob entry points move the frames over PCIe as bytes and widen them on the host; the float
    blobs they deliver must be bit-identical to the plain float transfer, for every chunking of the batch and for
    blobs that are not cache-line aligned."""
    tasks = ofdg.ParamStream(7).generate(n)
    monkeypatch.setenv("OFDG_TRANSPORT", "f32")
    gf = _gen(ofdg, 7, max_batch=64)
    monkeypatch.delenv("OFDG_TRANSPORT")
    gb = _gen(ofdg, 7, max_batch=64)
    P = 384 * 512
    outs = []
    for g in (gf, gb):
        g.upload_textures(textures8)
        raw = [np.full(n * c * P + 3, -7.0, np.float32) for c in (3, 3, 2)]
        views = [r[3:].reshape(n, c, 384, 512) for r, c in zip(raw, (3, 3, 2))]   # 12 bytes off any 64-byte boundary
        g.render_host(tasks, *views)
        assert all(np.all(r[:3] == -7.0) for r in raw)
        outs.append(views)
    assert gf.last_download_bytes() == n * 8 * P * 4
    assert gb.last_download_bytes() == n * (6 * P + 2 * P * 4)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    # the parameter-stream flavour goes through the same pipeline
    h = [np.empty((n, c, 384, 512), np.float32) for c in (3, 3, 2)]
    gb.generate_host(ofdg.ParamStream(7), n, *h)
    for a, b in zip(outs[1], h):
        assert np.array_equal(a, b)
    gf.close(); gb.close()


def test_host_pipeline_reports_a_bad_task_and_recovers(ofdg, textures8):
    """A descriptor the reference would reject (DataGenerator.cpp:1143) inside a later chunk: the call fails with the
    flattening error, nothing keeps running behind it, and the generator stays usable."""
    g = _gen(ofdg, 7, max_batch=64)
    g.upload_textures(textures8)
    good = ofdg.ParamStream(7).generate(20)
    arrs = good.arrays()
    first_fg = arrs["task_begin"][13] + 1
    arrs["blueprints"]["obj_type"][first_fg] = 77
    bad = ofdg.Tasks.from_arrays(arrs)
    with pytest.raises(ofdg.OfdgError, match="Bad object type"):
        g.render_host(bad)
    a = g.render_host(good)
    b = g.render_host(good)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    g.close()


@pytest.mark.parametrize("mode", [7, 9])
def test_extra_tops_backward_flow_ids_occlusion(ofdg, oracle, textures8, fields4, mode):
    """SURVEY 8 f4 extra tops: index images and computeFlowImage(inverse=true) (DataGenerator.cpp:740-818) against the
    oracle, plus this repository's occlusion mask against its restatement. The three standard blobs do not change."""
    import torch
    n = 6
    g = _gen(ofdg, mode)
    g.upload_textures(textures8)
    if mode == 9:
        g.set_fields(fields4)
    tasks = ofdg.ParamStream(mode, n_fields=4 if mode == 9 else 0).generate(n)
    dev = lambda c: torch.full((n, c, 384, 512), -5.0, device="cuda")
    i0, i1, fl = dev(3), dev(3), dev(2)
    g.render(tasks, i0, i1, fl)
    torch.cuda.synchronize()
    plain = [t.clone() for t in (i0, i1, fl)]
    bw, id0, id1, occ = dev(2), dev(1), dev(1), dev(1)
    g.set_extra_tops(flow_bw=bw, id0=id0, id1=id1, occlusion=occ)
    g.render(tasks, i0, i1, fl)
    torch.cuda.synchronize()
    for a, b in zip(plain, (i0, i1, fl)):
        assert torch.equal(a, b) or (mode == 9 and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)))
    cpu = oracle.render(tasks.struct(), textures8, mode=mode, fields=fields4 if mode == 9 else None, debug=True)
    assert np.array_equal(id0.cpu().numpy()[:, 0], cpu["id0"].astype(np.float32))
    assert np.array_equal(id1.cpu().numpy()[:, 0], cpu["id1"].astype(np.float32))
    d = np.abs(bw.cpu().numpy() - cpu["flow_bw"])
    assert np.nanmax(d) <= FLOW_TOL and np.array_equal(np.isnan(bw.cpu().numpy()), np.isnan(cpu["flow_bw"]))
    o_gpu, o_cpu = occ.cpu().numpy(), cpu["occlusion"]
    assert set(np.unique(o_gpu)) <= {0.0, 1.0}
    # the mask is a function of flow (compared to 1e-3 px, not bitwise): allow disagreement only where a flow component
    # sits within the tolerance of a rounding boundary
    assert (o_gpu != o_cpu).mean() <= 1e-5
    assert 0.001 < o_gpu.mean() < 0.6
    # numpy restatement of the definition from the GPU's own tops: exact
    f = fl.cpu().numpy(); a0 = id0.cpu().numpy()[:, 0]; a1 = id1.cpu().numpy()[:, 0]
    ys, xs = np.mgrid[0:384, 0:512].astype(np.float32)
    with np.errstate(invalid="ignore"):
        tx, ty = xs + f[:, 0], ys + f[:, 1]
        inside = (tx >= -0.5) & (tx < 511.5) & (ty >= -0.5) & (ty < 383.5)
        qx = np.where(inside, np.floor(tx + np.float32(0.5)), 0).astype(np.int64)
        qy = np.where(inside, np.floor(ty + np.float32(0.5)), 0).astype(np.int64)
    same = a1[np.arange(n)[:, None, None], qy, qx] == a0
    assert np.array_equal(o_gpu[:, 0], np.where(inside & same, 0.0, 1.0).astype(np.float32))
    g.set_extra_tops()   # off again: the registered tensors are left alone
    bw.fill_(-9.0)
    g.render(tasks, i0, i1, fl)
    torch.cuda.synchronize()
    assert float(bw.min()) == -9.0
    g.close()


def _mixed_pool(ofdg):
    """Textures of six sizes: every branch of Texture::getRandomizedCrop (DataGenerator.cpp:87-109) for 512 x 384 frames."""
    sizes = [(1024, 768),   # >= 2W x 2H: background crop + resize, foreground centre crop
             (300, 200),    # smaller than W x H: foreground resized up on both axes; background resized whole
             (640, 480),    # >= W x H but < 2W x 2H: foreground crop, background resized whole (growing)
             (1500, 500),   # wide and flat: background resized whole, shrinking along x (many taps), growing along y
             (400, 900),    # narrow and tall: foreground x grows / y shrinks 2.3x; background likewise
             (2600, 2000)]  # much larger than 2W x 2H: crop path with far-away origins
    return [ofdg.synth_textures(1, w, h, seed=20 + i)[0] for i, (w, h) in enumerate(sizes)]


def test_mixed_size_pool_views_and_backgrounds(ofdg, oracle):
    g = _gen(ofdg, 7)
    pool = _mixed_pool(ofdg)
    g.upload_textures(pool)
    for i, t in enumerate(pool):
        assert g.texture_size(i) == (t.shape[2], t.shape[1])
        assert np.array_equal(g.download_texture(i), t)
        assert np.array_equal(g.download_foreground_view(i), oracle.randomized_crop(t, 512, 384)), f"foreground view of texture {i}"
    # backgrounds: force every texture through the background path with the parameters the stream drew
    tasks = ofdg.ParamStream(7).generate(12)
    a = tasks.arrays()
    for t in range(12):
        a["blueprints"]["tex_id"][a["task_begin"][t]] = t % 6
    tasks = ofdg.Tasks.from_arrays(a)
    bg, need = g.debug_background(tasks)
    for t in range(12):
        b = a["blueprints"][a["task_begin"][t]]
        ref = oracle.randomized_crop(pool[t % 6], 1024, 768, float(b["tex_rot"]), float(b["tex_scale"]), int(b["tex_shift_x"]), int(b["tex_shift_y"]))
        x0, y0, x1, y1 = need[t]
        assert np.array_equal(bg[t][:, y0:y1 + 1, x0:x1 + 1], ref[:, y0:y1 + 1, x0:x1 + 1]), f"task {t} texture {t % 6}"
    g.close()


def test_mixed_size_pool_render_parity(ofdg, oracle):
    g = _gen(ofdg, 7)
    pool = _mixed_pool(ofdg)
    g.upload_textures(pool)
    tasks = ofdg.ParamStream(7).generate(8)
    a = tasks.arrays()
    used = set(int(v) % 6 for v in a["blueprints"]["tex_id"][a["blueprints"]["parent"] < 0])
    assert used == set(range(6)), "every pool texture must be in use"
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), pool, mode=7, debug=True)
    assert _compare(gpu, cpu) <= IMG_TOL
    g.close()


def _ellipse(obj_id, x, y, rx, ry, tex_id=3, rot=0.1, trans=(5.0, -3.0)):
    return dict(obj_id=obj_id, obj_type=1, init_rot=0.3, init_scale=0.0, init_trans_x=x, init_trans_y=y, rot=rot, scale=1.02,
                trans_x=trans[0], trans_y=trans[1], tex_id=tex_id, tex_rot=0.0, tex_scale=0.0, tex_shift_x=0, tex_shift_y=0,
                ellipse_scale_x=rx, ellipse_scale_y=ry, seg_begin=0, seg_count=0, comp_begin=0, comp_count=0, parent=-1,
                is_additive_component=0, do_warpfield_deformation=0, field_id=-1)


def test_edge_case_scenes(ofdg, oracle, textures8):
    """Hand-made scenes at the corners of the input space: a sample with no foreground at all, objects entirely outside the
    frame, objects hanging over every border (negative coordinates feed the carry-in of the tile rasteriser), one ellipse
    larger than the frame, a sliver thinner than a pixel, and a degenerate polygon with zero area."""
    base = ofdg.ParamStream(5).generate(1).arrays()
    bg = base["blueprints"][0]
    dt = base["blueprints"].dtype

    def rec(d):
        r = np.zeros(1, dt)
        for k, v in d.items():
            r[k] = v
        return r[0]

    scenes = [
        [],                                                                     # background only
        [_ellipse(10, -400.0, -300.0, 40, 30), _ellipse(11, 2000.0, 900.0, 50, 50)],  # everything off-frame
        [_ellipse(10, 0.0, 100.0, 60, 40), _ellipse(11, 511.0, 200.0, 70, 30), _ellipse(12, 250.0, 0.0, 30, 80),
         _ellipse(13, 300.0, 383.0, 90, 25), _ellipse(14, -20.0, -10.0, 45, 45)],      # over every border and the corner
        [_ellipse(10, 256.0, 192.0, 900, 700)],                                 # larger than the frame
        [_ellipse(10, 200.0, 150.0, 120, 0.2), _ellipse(11, 300.0, 250.0, 0.3, 90)],  # slivers
    ]
    bps, task_begin = [], [0]
    seg_type, seg_x, seg_y = [], [], []
    for sc in scenes:
        bps.append(bg)
        bps += [rec(d) for d in sc]
        task_begin.append(len(bps))
    # a degenerate polygon (all vertices on one line) next to a proper triangle that leaves the frame on the left
    bps.append(bg)
    for k, pts in enumerate([[(-40, -40), (0, 0), (40, 40), (80, 80)], [(-300, -60), (90, 0), (-300, 60)]]):
        d = _ellipse(10 + k, 260.0, 190.0, 0, 0)
        d.update(obj_type=2, seg_begin=len(seg_type), seg_count=len(pts))
        bps.append(rec(d))
        for (x, y) in pts:
            seg_type.append(1); seg_x.append(float(x)); seg_y.append(float(y))   # OFDG_SEG_LINE
    task_begin.append(len(bps))
    arrs = {"task_begin": np.array(task_begin, np.int32), "blueprints": np.array(bps, dt), "seg_type": np.array(seg_type, np.int32),
            "seg_x": np.array(seg_x, np.float32), "seg_y": np.array(seg_y, np.float32), "augment": None}
    tasks = ofdg.Tasks.from_arrays(arrs)
    g = _gen(ofdg, 5)
    g.upload_textures(textures8)
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), textures8, mode=5, debug=True)
    _compare(gpu, cpu)
    assert (gpu["id0"][0] == 1).all() and (gpu["id0"][1] == 1).all()          # nothing but background
    assert (gpu["id0"][3] == 10).mean() > 0.95                                # the giant ellipse hides it
    assert (gpu["id0"][2] != 1).any() and (gpu["masks"][4, :2] > 0).any()     # border objects and slivers leave a trace
    # the collinear polygon is at most a hairline once its vertices are snapped to 1/256 px; the triangle covers an area
    assert (gpu["masks"][5, 0, 0] > 0).sum() < 600 and (gpu["masks"][5, 1, 0] > 0).sum() > 2000
    g.close()


def test_argument_and_state_errors(ofdg, textures8):
    import torch
    g = _gen(ofdg, 7, max_batch=4)
    tasks = ofdg.ParamStream(7).generate(2)
    with pytest.raises(ofdg.OfdgError, match="no textures"):
        g.render_host(tasks)
    g.upload_textures(textures8)
    with pytest.raises(ofdg.OfdgError, match="max_batch"):
        g.render_host(ofdg.ParamStream(7).generate(5))
    with pytest.raises(ofdg.OfdgError, match="empty"):
        g.render_host(ofdg.Tasks())
    many = ofdg.ParamStream(7, fg_override=300).generate(1)
    with pytest.raises(ofdg.OfdgError, match="254"):
        g.render_host(many)
    a = tasks.arrays()
    a["blueprints"]["tex_scale"][0] = 1e-6          # a background crop thousands of times the prepared size
    with pytest.raises(ofdg.OfdgError, match="40x"):
        g.render_host(ofdg.Tasks.from_arrays(a))
    out = g.render_host(tasks)                       # the generator survives all of the above
    assert out[0].std() > 10
    g.close()
    with pytest.raises(ofdg.OfdgError, match="BAD MODE"):
        ofdg.Generator(device=0, mode=14)
    with pytest.raises(ofdg.OfdgError, match="multiple of 4"):
        ofdg.Generator(device=0, mode=1, width=510)


def test_pair_buffer_overflow_is_reported(ofdg, textures8, monkeypatch):
    """The (object, tile) pair buffers are sized from an upper bound, so the binning kernel cannot run out of room; if it
    ever did, the raster and shade kernels skip the batch and the call fails instead of returning blobs nobody rendered.
    OFDG_TEST_PAIR_CAP makes the kernels believe the buffers hold 8 pairs."""
    monkeypatch.setenv("OFDG_TEST_PAIR_CAP", "8")
    g = _gen(ofdg, 7, max_batch=4)
    monkeypatch.delenv("OFDG_TEST_PAIR_CAP")
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(7).generate(2)
    with pytest.raises(ofdg.OfdgError, match="pairs"):
        g.render_host(tasks)
    g.close()
    g = _gen(ofdg, 7, max_batch=4)   # without the limit the same batch renders
    g.upload_textures(textures8)
    assert g.render_host(tasks)[0].std() > 10
    g.close()


@pytest.mark.parametrize("mode", [7, 9])
def test_prepared_batches_into_host_blobs(ofdg, oracle, textures8, fields4, mode):
    """ofdg_render_prepared_host (the layer's Forward_cpu): chunked windows over a resident scene, uint8 transport,
    mode 9's warped masks made once -- identical to rendering the prepared batch into device blobs."""
    import torch
    n = 12
    g = _gen(ofdg, mode)
    g.upload_textures(textures8)
    if mode == 9:
        g.set_fields(fields4)
    tasks = ofdg.ParamStream(mode, n_fields=4 if mode == 9 else 0).generate(n)
    p = g.prepare(tasks)
    dev = [torch.empty((n, c, 384, 512), device="cuda") for c in (3, 3, 2)]
    g.render_prepared(p, *dev)
    torch.cuda.synchronize()
    host = g.render_prepared_host(p)
    for d, h in zip(dev, host):
        assert np.array_equal(d.cpu().numpy(), h, equal_nan=True)
    assert g.last_download_bytes() == n * (6 + 8) * 384 * 512
    g.close()


@pytest.mark.parametrize("W,H", [(260, 150), (132, 84)])
def test_frame_sizes_off_the_tile_grid(ofdg, oracle, textures8, W, H):
    """Runtime frame sizes that are no multiple of the 128 x 8 render tile or the 32 x 32 preparation tile (the reference
    fixes 512 x 384 at compile time, DataGenerator.h:55-56): partial tiles at the right and bottom borders."""
    g = ofdg.Generator(device=0, width=W, height=H, mode=7, max_batch=3)
    g.upload_textures(textures8)
    tasks = ofdg.ParamStream(7, W, H).generate(3)
    gpu = g.render_debug(tasks)
    cpu = oracle.render(tasks.struct(), textures8, W=W, H=H, mode=7, debug=True)
    _compare(gpu, cpu)
    host = g.render_host(tasks)          # chunked uint8 transport with planes that are not 64-byte multiples
    assert np.array_equal(host[0], gpu["img0"]) and np.array_equal(host[2], gpu["flow"])
    g.close()
