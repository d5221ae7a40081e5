"""Host parameter stream (ObjectParametersGenerator restated): per-engine draw
bookkeeping for every mode branch, value ranges, determinism, sharding offsets, resume."""
import numpy as np
import pytest

SLOTS = ["BgTexID", "BgInitRot", "BgInitTransX", "BgInitTransY", "BgRotTrigger", "BgRot", "BgTransX", "BgTransY",
         "BgScaleTrigger", "BgInitScale", "BgScale", "NumberOfFgObjects", "ObjType", "ObjTexID", "ObjInitTransX",
         "ObjInitTransY", "ObjTransX", "ObjTransY", "ObjInitRot", "ObjRotTrigger", "ObjRot", "ObjInitScale",
         "ObjScaleTrigger", "ObjScale", "ObjTexShiftX", "ObjTexShiftY", "ObjTexRot", "ObjTexZoom", "ElliObj_ScaleX",
         "ElliObj_ScaleY", "PolyObj_spokes", "PolyObj_dphi", "PolyObj_r", "PolyObj_ScaleX", "PolyObj_ScaleY",
         "PolyObj_CurveTrigger", "CompObjInitTransX", "CompObjInitTransY", "CompObiNumberOfComponents",
         "ComponentIsAdditive", "ComponentOffset", "ObjIsExtraThin", "ObjDeformsNonrigidly", "GenericUniform",
         "GenericTrigger"]
S = {n: i for i, n in enumerate(SLOTS)}
NEVER = ["ObjInitScale", "ObjTexShiftX", "ObjTexShiftY", "ObjTexRot", "ObjTexZoom", "GenericUniform"]


def test_slot_names_follow_the_declaration_order(ofdg):
    assert ofdg.slot_names() == SLOTS  # DataGenerator.h:524-587; slot index == seed


@pytest.mark.parametrize("mode", range(1, 14))
def test_draw_bookkeeping(ofdg, mode):
    n = 40
    ps = ofdg.ParamStream(mode)
    a = ps.generate(n).arrays()
    d = ps.draws()
    bp, tb = a["blueprints"], a["task_begin"]
    top = bp[bp["parent"] < 0]
    bg = top[top["obj_id"] == 1]
    fg = top[top["obj_id"] >= 10]
    comps = bp[bp["parent"] >= 0]
    assert len(bg) == n
    # background: one draw each, BgRot / BgScale only when their trigger fired (DataGenerator.cpp:2122-2136)
    for s in ("BgTexID", "BgInitRot", "BgInitTransX", "BgInitTransY", "BgRotTrigger", "BgTransX", "BgTransY",
              "BgScaleTrigger", "BgInitScale", "NumberOfFgObjects"):
        assert d[S[s]] == n, s
    # (a drawn value can still be exactly 0: out-of-range Gaussians fall back to the midpoint)
    assert int((bg["rot"] != 0).sum()) <= d[S["BgRot"]] <= n and int((bg["scale"] != 1).sum()) <= d[S["BgScale"]] <= n
    if mode in (1, 2, 3, 8):
        assert d[S["BgRot"]] == n and np.all(bg["rot"] == 0) and np.all(bg["scale"] == 1)  # Trigger(0,0,1) always fires, range (0,0)
    for s in NEVER:
        assert d[S[s]] == 0, s
    # the foreground common prefix is drawn once per object and once per component
    # ("outline" composites copy their second part from the first without drawing: DataGenerator.cpp:2520)
    n_outline = int((fg["obj_type"] == ofdg.OBJ_COMPOSITE).sum()) - d[S["CompObiNumberOfComponents"]]
    assert n_outline >= 0 and (n_outline == 0 or mode in (7, 9, 10, 11, 12, 13))
    n_obj = len(fg) + len(comps) - n_outline
    for s in ("ObjInitTransX", "ObjInitTransY", "ObjRotTrigger", "ObjScaleTrigger", "ObjTransX", "ObjTransY", "ObjTexID"):
        assert d[S[s]] == n_obj, s
    assert d[S["ObjType"]] >= n_obj  # components redraw while they draw Composite
    assert d[S["ObjDeformsNonrigidly"]] == n + (n_obj if mode == 9 else 0)
    # object counts: (int)Uniform(16,24) -> 16..23
    per_task = np.array([int(((bp[tb[t]:tb[t + 1]]["parent"] < 0) & (bp[tb[t]:tb[t + 1]]["obj_id"] >= 10)).sum()) for t in range(n)])
    assert per_task.min() >= 16 and per_task.max() <= 23
    # ids: background 1, k-th object 10+k, components 0 (data_generation_layer.cpp:201, 210)
    for t in range(n):
        b = bp[tb[t]:tb[t + 1]]
        assert b[0]["obj_id"] == 1
        tops = b[(b["parent"] < 0)][1:]
        assert list(tops["obj_id"]) == list(range(10, 10 + len(tops)))
    assert np.all(comps["obj_id"] == 0)
    # mode-specific shapes
    types = set(np.unique(fg["obj_type"]))
    allowed = {1: {2}, 2: {2}, 3: {1}, 4: {1, 2}, 5: {1, 2}, 8: {1, 2}}.get(mode, {1, 2, 3})
    assert types <= allowed
    assert set(np.unique(comps["obj_type"])) <= {1, 2}
    if mode < 6 or mode == 8:
        assert len(comps) == 0 and d[S["CompObiNumberOfComponents"]] == 0
    if mode in (1, 2):
        assert d[S["PolyObj_CurveTrigger"]] == 0 and not (a["seg_type"] == ofdg.SEG_CURVE3).any()
    if mode == 1:
        assert d[S["PolyObj_spokes"]] == 0 and np.all(fg["seg_count"] == 4)
    if mode not in (7, 9, 10, 11, 12, 13):
        assert d[S["ObjIsExtraThin"]] == 0 and d[S["GenericTrigger"]] == 0
    else:
        assert d[S["ObjIsExtraThin"]] == len(fg)  # once per top-level object, never for components
        assert d[S["ComponentIsAdditive"]] == d[S["ComponentOffset"]] // 2


def test_value_ranges_mode7(ofdg):
    a = ofdg.ParamStream(7).generate(200).arrays()
    bp = a["blueprints"]
    bg = bp[bp["obj_id"] == 1]
    fg = bp[(bp["parent"] < 0) & (bp["obj_id"] >= 10)]
    assert np.abs(bg["rot"]).max() <= 10 * np.pi / 180 + 1e-6
    assert bg["scale"].min() >= 0.93 - 1e-6 and bg["scale"].max() <= 1.07 + 1e-6
    assert 0.8 <= bg["tex_scale"].min() and bg["tex_scale"].max() <= 1.2
    assert set(np.unique(bg["tex_shift_x"])) <= {0, 512} and set(np.unique(bg["tex_shift_y"])) <= {0, 384}
    assert np.abs(bg["tex_rot"]).max() <= np.pi + 1e-6
    assert (bg["rot"] != 0).mean() == pytest.approx(0.3, abs=0.12) and (bg["scale"] != 1).mean() == pytest.approx(0.6, abs=0.12)
    assert fg["init_trans_x"].min() >= -306 and fg["init_trans_x"].max() <= 818
    assert fg["init_trans_y"].min() >= -242 and fg["init_trans_y"].max() <= 626
    assert np.abs(fg["trans_x"]).max() <= 120 and np.abs(fg["rot"]).max() <= 30 * np.pi / 180 + 1e-6
    assert fg["scale"].min() >= 0.8 - 1e-6 and fg["scale"].max() <= 1.2 + 1e-6
    assert fg["tex_id"].min() >= 0
    ell = fg[fg["obj_type"] == ofdg.OBJ_ELLIPSE]
    assert ell["ellipse_scale_y"].min() >= 25 and ell["ellipse_scale_y"].max() <= 100
    assert ell["ellipse_scale_x"].min() >= 25 * 0.05 - 1e-4  # extra thin: x radius * 0.05
    comp = fg[fg["obj_type"] == ofdg.OBJ_COMPOSITE]
    assert comp["comp_count"].min() >= 1 and comp["comp_count"].max() <= 7
    for c in comp:
        parts = bp[c["comp_begin"]:c["comp_begin"] + c["comp_count"]]
        assert parts[0]["is_additive_component"] == 1
        assert np.all(parts["rot"] == c["rot"]) and np.all(parts["scale"] == c["scale"]) and np.all(parts["trans_x"] == c["trans_x"])
        assert parts[0]["init_trans_x"] == c["init_trans_x"] and parts[0]["init_rot"] == c["init_rot"]


def test_deterministic_and_resumable(ofdg):
    a = ofdg.ParamStream(7).generate(12).arrays()
    b = ofdg.ParamStream(7).generate(12).arrays()
    for k in a:
        assert np.array_equal(a[k], b[k])
    ps = ofdg.ParamStream(7)
    ps.skip(8)
    assert ps.tasks_generated() == 8
    c = ps.generate(4).arrays()
    tail = ofdg.select_tasks(a, range(8, 12))
    for k in c:
        assert np.array_equal(c[k], tail[k]), k


def test_seed_offset_shards_differ(ofdg):
    a = ofdg.ParamStream(7, seed_offset=0).generate(4).arrays()
    b = ofdg.ParamStream(7, seed_offset=45).generate(4).arrays()
    assert not np.array_equal(a["blueprints"]["tex_id"][:10], b["blueprints"]["tex_id"][:10])


def test_output_size_scales_the_tables(ofdg):
    a = ofdg.ParamStream(7, 1024, 768).generate(50).arrays()["blueprints"]
    fg = a[(a["parent"] < 0) & (a["obj_id"] >= 10)]
    assert fg["init_trans_x"].max() > 900  # Uniform(-W/2-50, 3W/2+50) follows the runtime W
    assert set(np.unique(a[a["obj_id"] == 1]["tex_shift_x"])) <= {0, 1024}


def test_mode9_field_ids(ofdg):
    ps = ofdg.ParamStream(9, n_fields=5)
    a = ps.generate(30).arrays()["blueprints"]
    flagged = a[(a["do_warpfield_deformation"] != 0)]
    assert len(flagged) > 0
    assert np.all(flagged["field_id"] >= 0) and np.all(flagged["field_id"] < 5)
    tops = flagged[flagged["parent"] < 0]["field_id"]
    # every crop is served three times before the next one (WarpFields.cpp:516-538)
    assert list(tops[:9]) == [0, 0, 0, 1, 1, 1, 2, 2, 2]
    comps = a[(a["parent"] >= 0)]
    for c in comps:
        assert c["field_id"] == a[c["parent"]]["field_id"] and c["do_warpfield_deformation"] == a[c["parent"]]["do_warpfield_deformation"]


def test_bad_mode_rejected(ofdg):
    with pytest.raises(ofdg.OfdgError, match="BAD MODE"):
        ofdg.ParamStream(14)


def test_augmentation_leaves_the_reference_stream_alone(ofdg, oracle, textures8):
    a = ofdg.ParamStream(7).generate(5).arrays()
    ps = ofdg.ParamStream(7)
    ps.enable_augmentation(True)
    b = ps.generate(5).arrays()
    for k in ("task_begin", "blueprints", "seg_type", "seg_x", "seg_y"):
        assert np.array_equal(a[k], b[k]), k
    assert a["augment"] is None and len(b["augment"]) == 5
    aug = b["augment"]
    assert np.all((aug["gain"] >= 0.8) & (aug["gain"] <= 1.2)) and np.all(np.abs(aug["brightness"]) <= 20)
    assert np.all((aug["contrast"] >= 0.7) & (aug["contrast"] <= 1.3)) and np.all((aug["noise_sigma"] >= 0) & (aug["noise_sigma"] <= 10))


def test_augmentation_noise_statistics(ofdg, oracle, textures8):
    """The Irwin-Hall/Philox noise of the augmentation spec is ~N(0, sigma) and differs per frame/channel."""
    t = ofdg.ParamStream(1).generate(1)
    arrs = t.arrays()
    aug = np.zeros(1, ofdg.AUGMENT_DTYPE)
    aug["enabled"], aug["gain"], aug["contrast"], aug["noise_sigma"], aug["noise_seed"] = 1, 1.0, 1.0, 5.0, (123, 456)
    arrs["augment"] = aug
    s, keep = ofdg.struct_from_arrays(arrs)
    noisy = oracle.render(s, textures8, mode=1, n_threads=1)
    arrs["augment"] = None
    s2, keep2 = ofdg.struct_from_arrays(arrs)
    clean = oracle.render(s2, textures8, mode=1, n_threads=1)
    d = (noisy["img0"] - clean["img0"])[0]
    inner = (clean["img0"][0] > 30) & (clean["img0"][0] < 225)  # away from the clamp
    assert abs(d[inner].mean()) < 0.05 and abs(d[inner].std() - 5.0) < 0.1
    d1 = (noisy["img1"] - clean["img1"])[0]
    assert np.corrcoef(d[0].ravel()[:50000], d[1].ravel()[:50000])[0, 1] < 0.02  # channels independent
    assert not np.array_equal(d[0], d1[0])
    assert np.array_equal(noisy["flow"], clean["flow"])


@pytest.mark.parametrize("mode", [1, 7, 9])
def test_lookahead_threads_leave_the_stream_unchanged(ofdg, mode):
    """ofdg_params_set_threads: helper threads produce the engines' values ahead of the sequential walk. The k-th value of an
    engine depends on its seed and k only, so blueprints, segments, per-engine draw counts and resume points are the same
    with and without them -- for batches of changing size, single tasks and a skip in between."""
    def run(threads):
        ps = ofdg.ParamStream(mode, n_fields=40 if mode == 9 else 0)
        if threads:
            ps.set_threads(threads)
        out = []
        for n in (64, 64, 3, 64, 17, 64, 64):
            a = ps.generate(n).arrays()
            out.append((a["blueprints"].tobytes(), a["seg_type"].tobytes(), a["seg_x"].tobytes(), a["seg_y"].tobytes()))
        ps.skip(5)
        a = ps.generate(64).arrays()
        out.append((a["blueprints"].tobytes(), a["seg_x"].tobytes()))
        return out, ps.draws(), ps.tasks_generated()
    base = run(0)
    for threads in (1, 3):
        got = run(threads)
        assert got[1] == base[1] and got[2] == base[2]
        for k, (x, y) in enumerate(zip(base[0], got[0])):
            assert x == y, f"batch {k} differs with {threads} look-ahead threads"
