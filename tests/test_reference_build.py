"""The REFERENCE'S OWN code as the checker (oracle/_ref, built by oracle/ref_build.sh from the untouched sources under
/root/reference against the stand-in AGG / CImg / Caffe headers of oracle/shim).

What this pins, on the CPU:
  * the product's host parameter stream (csrc/host/params.cpp) against the reference's ObjectParametersGenerator driven by
    the commission loop of its layer -- byte for byte, every mode, several per-GPU seed offsets;
  * the oracle restatement (oracle/oracle.cpp, oracle/warpfields.cpp) against the reference's own DataGenerator /
    RenderCore / MovingObject* / WarpFields classes -- masks, index images, frames and flows bit for bit (a5-a20, f3);
  * the blob contract of the reference's Caffe layer end to end (prototxt parameters -> texture list -> Forward_cpu).
What it cannot pin: the arithmetic inside AGG 2.4 and CImg, which both sides take from a restatement (oracle.cpp inlines
it, oracle/shim wraps it in the libraries' own interfaces) because neither library is available offline.
The GPU suite compares the sm_100a path with the same library (test_gpu_matches_the_reference_build)."""
import numpy as np
import pytest


def _same_floats(a, b):
    return bool((((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))).all())


@pytest.mark.parametrize("mode", list(range(1, 14)))
def test_parameter_stream_equals_the_reference_generator(ofdg, refimpl, mode):
    """ofdg_params_generate == ObjectParametersGenerator (DataGenerator.cpp:1353-2835) under load_batch's commission loop
    (data_generation_layer.cpp:197-214): every blueprint field, every polygon segment, 256 tasks, seed offsets 0 / 45 / 315.
    field_id is this repository's own bookkeeping (the reference takes the next crop of its queue) and is checked through
    test_mode9_* below instead."""
    for off in (0, 45, 315):
        mine = ofdg.ParamStream(mode, seed_offset=off).generate(256).arrays()
        ref = refimpl.ParamStream(mode, off).generate(256)
        assert np.array_equal(mine["task_begin"], ref["task_begin"])
        assert np.array_equal(mine["seg_type"], ref["seg_type"])
        assert mine["seg_x"].tobytes() == ref["seg_x"].tobytes() and mine["seg_y"].tobytes() == ref["seg_y"].tobytes()
        a, b = mine["blueprints"].copy(), ref["blueprints"].copy()
        a["field_id"] = -1
        b["field_id"] = -1
        for name in a.dtype.names:
            assert a[name].tobytes() == b[name].tobytes(), (mode, off, name)


def test_parameter_stream_batches_concatenate(ofdg, refimpl):
    """Three calls of 5 tasks draw what one call of 15 draws (the layer tops up its queue in uneven portions)."""
    r = refimpl.ParamStream(7, 0)
    for _ in range(3):
        ref = r.generate(5)
    mine = ofdg.ParamStream(7).generate(15).arrays()
    assert mine["blueprints"]["rot"].tobytes() == ref["blueprints"]["rot"].tobytes()
    assert mine["seg_x"].tobytes() == ref["seg_x"].tobytes()


def _compare_renders(a, b):
    for k in ("masks", "id0", "id1", "frames8"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("img0", "img1"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("flow", "flow_bw"):
        assert _same_floats(a[k], b[k]), k


@pytest.mark.parametrize("mode,use_aa", [(1, True), (2, True), (3, False), (4, True), (6, True), (7, True), (7, False), (11, True)])
def test_restatement_equals_the_reference_render(ofdg, oracle, refimpl, textures8, mode, use_aa):
    """oracle.cpp against the reference's own Process_TaskBucket and the objects it builds (DataGenerator.cpp:1065-1254 and
    everything it calls): four masks per object, both index images, both frames, forward and backward flow -- bit for bit."""
    tasks = ofdg.ParamStream(mode).generate(2)
    s = tasks.struct()
    cpu = oracle.render(s, textures8, mode=mode, use_aa=use_aa, debug=True, n_threads=2)
    ref = refimpl.Generator(mode, use_aa=use_aa, textures=textures8).render(s, debug=True)
    _compare_renders(cpu, ref)


@pytest.fixture(scope="module")
def fields6(oracle):
    return oracle.generate_fields(seed=5, n_fields=6)


def test_field_producer_restatement_equals_the_reference(oracle, refimpl, fields6):
    """oracle/warpfields.cpp against the reference's DisplacementComposer + FlowField::init_from_DisplacementComposer +
    clamp_near_zeros + get_crop (WarpFields.cpp:337-455, 603-634), same seeded displacer scene: bit for bit, NaNs included."""
    ref = refimpl.generate_fields(5, 3)
    assert _same_floats(fields6[:3], ref)


def test_mode9_restatement_equals_the_reference_render(ofdg, oracle, refimpl, textures8, fields6):
    """Mode 9 (non-rigid motion) with an injected field pool: tasks chosen so that a deformed background, deformed
    composites and deformed plain objects all occur. field_policy 0 feeds the reference's crop queue by the batch's field
    ids; field_policy 1 hands it the pool in order and lets CropGenerator::get_crop's own reuse rule pick (every crop is
    served three times, WarpFields.cpp:516-538) -- which is what the product's id assignment has to reproduce."""
    stream = ofdg.ParamStream(9, n_fields=6)
    tasks = stream.generate(24)
    a = tasks.arrays()
    tb, bp = a["task_begin"], a["blueprints"]
    bg_deformed = [t for t in range(24) if bp[tb[t]]["do_warpfield_deformation"]]
    comp_deformed = [t for t in range(24) if any(r["obj_type"] == 3 and r["do_warpfield_deformation"] for r in bp[tb[t]:tb[t + 1]])]
    assert bg_deformed and comp_deformed
    # a prefix keeps the reference's own crop cursor (policy 1) in step with the stream's ids
    n = min(max(bg_deformed[0], comp_deformed[0]) + 1, 6)
    sel = tasks.select(range(n))
    assert any(t < n for t in bg_deformed) or any(t < n for t in comp_deformed)
    s = sel.struct()
    cpu = oracle.render(s, textures8, mode=9, debug=True, n_threads=4, fields=fields6)
    g = refimpl.Generator(9, textures=textures8, fields=fields6)
    _compare_renders(cpu, g.render(s, debug=True, field_policy=0))
    _compare_renders(cpu, g.render(s, debug=True, field_policy=1))


def test_randomized_crop_both_branches(ofdg, oracle, refimpl):
    """Texture::getRandomizedCrop (DataGenerator.cpp:87-109) through the reference's own statement chain, for textures that
    take the crop branch and for smaller ones that are resized whole; foreground defaults and background arguments."""
    pool = [ofdg.synth_textures(1, w, h, seed=10 + i)[0] for i, (w, h) in enumerate([(1024, 768), (1100, 900), (640, 480), (300, 200), (1500, 700)])]
    g = refimpl.Generator(1, textures=pool)
    for i, t in enumerate(pool):
        assert np.array_equal(g.randomized_crop(i, 512, 384), oracle.randomized_crop(t, 512, 384)), ("foreground", i)
        for angle, zoom, sx, sy in [(0.7, 0.83, 0, 0), (-2.9, 1.17, 512, 384), (3.1, 1.0, 512, 0), (0.0, 0.9, 0, 384)]:
            assert np.array_equal(g.randomized_crop(i, 1024, 768, angle, zoom, sx, sy),
                                  oracle.randomized_crop(t, 1024, 768, angle, zoom, sx, sy)), ("background", i, angle, zoom)


def test_reference_layer_end_to_end(ofdg, oracle, refimpl, tmp_path):
    """The reference's DataGenerationLayer created through its own REGISTER_LAYER_CLASS entry, fed the prototxt parameters of
    example-prototxt/train.prototxt (mode 7) and a texture list of PPM files: LayerSetUp shapes the tops {N,3,384,512} x2 and
    {N,2,384,512} (data_generation_layer.cpp:128-130); with one first-level worker the k-th sample of Forward is the k-th
    task of the parameter stream, and equals the oracle's render of the product's parameter stream with the same pool."""
    tex = ofdg.synth_textures(3, 1024, 768, seed=21)
    lst = refimpl.write_ppm_pool(str(tmp_path / "pool"), tex)
    layer = refimpl.Layer(7, lst, batch=2, prefetch=1, first_level_threads=1, second_level_threads=1)
    assert layer.top_shape(0) == (2, 3, 384, 512) and layer.top_shape(1) == (2, 3, 384, 512) and layer.top_shape(2) == (2, 2, 384, 512)
    first = layer.forward()
    second = layer.forward()
    layer.close()
    tasks = ofdg.ParamStream(7).generate(4)
    cpu = oracle.render(tasks.struct(), tex, mode=7, n_threads=4)
    for k, name in enumerate(("img0", "img1", "flow")):
        assert np.array_equal(first[k], cpu[name][:2]), name
        assert np.array_equal(second[k], cpu[name][2:]), name


def test_reference_texture_list_semantics(ofdg, refimpl, tmp_path):
    """TextureCollection (DataGenerator.cpp:117-149): R and B planes are swapped after loading, and a last line without a
    trailing newline is dropped (getline sets eof, the loop breaks before loading it). The product's list reader mirrors both."""
    tex = ofdg.synth_textures(3, 64, 48, seed=2)
    lst = refimpl.write_ppm_pool(str(tmp_path / "pool"), tex)
    g = refimpl.Generator(1)
    g.load_list(lst)
    assert g.texture_count() == 3
    assert np.array_equal(g.texture(1), tex[1])  # planes B,G,R in memory, R,G,B on disk
    text = open(lst).read()
    open(lst, "w").write(text.rstrip("\n"))
    g.load_list(lst)
    assert g.texture_count() == 2


def test_golden_vectors_are_reference_outputs(ofdg, refimpl):
    """tests/golden/render_golden.npz equals what the reference build renders (the file pins reference outputs, not only the restatement)."""
    import test_golden as tg
    for mode, n in [(1, 2), (12, 2)]:
        tex = ofdg.synth_textures(8, 1024, 768, seed=1)
        tasks = ofdg.ParamStream(mode).generate(n)
        out = refimpl.Generator(mode, textures=tex).render(tasks.struct(), debug=True)
        tg._check(mode, tg.mg.summarise(out), 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [5, 7, 9])
def test_gpu_matches_the_reference_build(ofdg, refimpl, mode):
    """The sm_100a path through the C ABI against the reference's own code: masks and index images bit-exact, uint8 frames
    within 1 LSB, flow within 1e-3 px (BASELINE.json north_star tolerances)."""
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    fields = None
    g = ofdg.Generator(device=0, mode=mode, max_batch=4)
    g.upload_textures(tex)
    if mode == 9:
        fields = g.generate_fields(seed=3, n=6)
    tasks = ofdg.ParamStream(mode, n_fields=0 if fields is None else 6).generate(4)
    gpu = g.render_debug(tasks)
    ref = refimpl.Generator(mode, textures=tex, fields=fields).render(tasks.struct(), debug=True)
    assert np.array_equal(gpu["masks"], ref["masks"])
    assert np.array_equal(gpu["id0"], ref["id0"]) and np.array_equal(gpu["id1"], ref["id1"])
    assert np.abs(gpu["frames8"].astype(np.int16) - ref["frames8"].astype(np.int16)).max() <= 1  # tolerance: 1 LSB
    assert np.abs(gpu["img0"] - ref["img0"]).max() <= 1 and np.abs(gpu["img1"] - ref["img1"]).max() <= 1
    both = ~(np.isnan(gpu["flow"]) | np.isnan(ref["flow"]))
    assert np.array_equal(np.isnan(gpu["flow"]), np.isnan(ref["flow"]))
    assert np.abs(gpu["flow"][both] - ref["flow"][both]).max() <= 1e-3  # tolerance: 1e-3 px
    g.close()
