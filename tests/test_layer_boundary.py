"""Caffe-layer boundary: prototxt parsing (the reference's example file verbatim), blob contract."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_example_prototxt_parses_verbatim(ofdg):
    text = open(os.path.join(GOLDEN, "train.prototxt")).read()  # copy of example-prototxt/train.prototxt
    p = ofdg.parse_prototxt(text)
    assert p == {"batch_size": 8, "prefetch": 40, "mode": 7, "first_level_threads": 8, "second_level_threads": 3,
                 "use_antialiasing": True, "top_size": 3, "type": "DataGeneration",
                 "texture_dbases": "/misc/lmbraid18/mayern/CLUSTER/resources/random-textures-1000/database.txt"}


def test_prototxt_defaults_and_errors(ofdg):
    p = ofdg.parse_prototxt('layer { type: "DataGeneration" top: "a" data_param { batch_size: 2 } data_generation_param { texture_dbases: "x" } }')
    assert p["mode"] == 1 and p["first_level_threads"] == 16 and p["second_level_threads"] == 1 and p["use_antialiasing"] is True
    with pytest.raises(ofdg.OfdgError, match="unknown"):
        ofdg.parse_prototxt('layer { type: "DataGeneration" data_generation_param { modee: 3 } }')
    with pytest.raises(ofdg.OfdgError):
        ofdg.parse_prototxt('layer { type: "DataGeneration" ')
    with pytest.raises(ofdg.OfdgError, match="takes a value"):  # used to dereference a null token
        ofdg.parse_prototxt('layer { type: "DataGeneration" top { } }')
    with pytest.raises(ofdg.OfdgError, match="DataGeneration"):
        ofdg.DataGenerationLayer('layer { type: "Data" top: "a" }')


def test_layer_registry(ofdg):
    """REGISTER_LAYER_CLASS(DataGeneration) (data_generation_layer.cpp:298-299): the type string is in the shim's
    LayerRegistry, and layers are built through LayerRegistry<float>::CreateLayer the way Net::Init builds them."""
    assert ofdg.DataGenerationLayer.registered_types() == ["DataGeneration"]
    with pytest.raises(ofdg.OfdgError, match="Unknown layer type: Convolution"):
        ofdg.DataGenerationLayer('layer { type: "Convolution" top: "a" }')


@pytest.mark.gpu
def test_layer_blob_contract(ofdg, oracle):
    import torch
    text = open(os.path.join(GOLDEN, "train.prototxt")).read()
    layer = ofdg.DataGenerationLayer(text, texture_db="synthetic:8:1")
    assert layer.type() == "DataGeneration"
    layer.LayerSetUp()
    assert layer.top_shape(0) == (8, 3, 384, 512) and layer.top_shape(1) == (8, 3, 384, 512) and layer.top_shape(2) == (8, 2, 384, 512)
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    ps = ofdg.ParamStream(7)
    for it in range(3):  # consecutive batches follow the commission order of the parameter stream
        if it % 2:
            layer.Forward_cpu()
        else:
            layer.Forward_gpu()
        got = [layer.top_cpu(i) for i in range(3)]
        tasks = ps.generate(8)
        ref = oracle.render(tasks.struct(), tex, mode=7)
        assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1
        assert np.abs(got[2] - ref["flow"]).max() <= 1e-3
        assert got[0].min() >= 0 and got[0].max() <= 255 and got[0].dtype == np.float32
    layer.close()


@pytest.mark.gpu
def test_layer_texture_list_of_mixed_files(ofdg, oracle, tmp_path):
    """texture_dbases as a list file (TextureCollection ctor, DataGenerator.cpp:117-149): a PPM, a PNG and a BMP of three
    different sizes, with the reference's R<->B swap."""
    import struct
    tex = [ofdg.synth_textures(1, w, h, seed=9 + i)[0] for i, (w, h) in enumerate([(1024, 768), (700, 500), (320, 240)])]  # planes = R,G,B on disk
    rgb = [np.ascontiguousarray(t.transpose(1, 2, 0)) for t in tex]
    (tmp_path / "t0.ppm").write_bytes(b"P6\n# comment\n1024 768\n255\n" + rgb[0].tobytes())
    (tmp_path / "t1.png").write_bytes(_png_bytes(rgb[1]))
    rows = b"".join(rgb[2][y, :, ::-1].tobytes() for y in range(239, -1, -1))
    (tmp_path / "t2.bmp").write_bytes(b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) +
                                      struct.pack("<IiiHHIIiiII", 40, 320, 240, 1, 24, 0, len(rows), 2835, 2835, 0, 0) + rows)
    (tmp_path / "db.txt").write_text("".join(str(tmp_path / n) + "\n" for n in ("t0.ppm", "t1.png", "t2.bmp")))
    layer = ofdg.DataGenerationLayer('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 4 prefetch: 2 } '
                                     'data_generation_param { mode: 5 texture_dbases: "%s" } }' % (tmp_path / "db.txt"))
    layer.LayerSetUp()
    layer.Forward_gpu()
    got = [layer.top_cpu(i) for i in range(3)]
    tasks = ofdg.ParamStream(5).generate(4)
    ref = oracle.render(tasks.struct(), [t[::-1].copy() for t in tex], mode=5)  # planes held as B,G,R
    assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1
    assert np.abs(got[2] - ref["flow"]).max() <= 1e-3
    layer.close()


@pytest.mark.gpu
def test_layer_device_params_mode(ofdg):
    """Extension: device_params: true -> Philox production mode through the same layer surface."""
    proto = ('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 4 prefetch: 2 } '
             'data_generation_param { mode: 7 texture_dbases: "synthetic:8:1" device_params: true seed: 42 } }')
    layer = ofdg.DataGenerationLayer(proto)
    layer.LayerSetUp()
    layer.Forward_gpu()
    a = [layer.top_cpu(i) for i in range(3)]
    layer.Forward_gpu()
    b = [layer.top_cpu(i) for i in range(3)]
    assert a[0].shape == (4, 3, 384, 512) and a[2].shape == (4, 2, 384, 512)
    assert a[0].std() > 10 and not np.array_equal(a[0], b[0])
    layer.close()
    # a second layer with the same seed replays the same batches
    layer2 = ofdg.DataGenerationLayer(proto)
    layer2.LayerSetUp()
    layer2.Forward_gpu()
    assert np.array_equal(layer2.top_cpu(0), a[0])
    layer2.close()


@pytest.mark.gpu
def test_layer_extra_tops(ofdg, oracle):
    """Seven tops: the reference's three, then backward flow, occlusion and the two index images."""
    proto = ('layer { type: "DataGeneration" top: "img0" top: "img1" top: "flow" top: "flow_bw" top: "occ" top: "id0" top: "id1" '
             'data_param { batch_size: 3 prefetch: 2 } data_generation_param { mode: 5 texture_dbases: "synthetic:8:1" } }')
    layer = ofdg.DataGenerationLayer(proto)
    layer.LayerSetUp()
    assert [layer.top_shape(i)[1] for i in range(7)] == [3, 3, 2, 2, 1, 1, 1]
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    ps = ofdg.ParamStream(5)
    for it in range(2):
        layer.Forward_cpu() if it else layer.Forward_gpu()
        got = [layer.top_cpu(i) for i in range(7)]
        ref = oracle.render(ps.generate(3).struct(), tex, mode=5, debug=True)
        assert np.abs(got[2] - ref["flow"]).max() <= 1e-3 and np.abs(got[3] - ref["flow_bw"]).max() <= 1e-3
        assert np.array_equal(got[5][:, 0], ref["id0"].astype(np.float32)) and np.array_equal(got[6][:, 0], ref["id1"].astype(np.float32))
        assert (got[4] != ref["occlusion"]).mean() <= 1e-5
    layer.close()
    with pytest.raises(ofdg.OfdgError, match="at most 7"):
        bad = ofdg.DataGenerationLayer(proto.replace('top: "id1"', 'top: "id1" top: "x"'))
        bad.LayerSetUp()


def _png_bytes(rgb, filter_types=(0, 1, 2, 3, 4), alpha=False, palette=False, interlace=False):
    """Minimal PNG writer (zlib + struct) exercising every scanline filter; interlace: Adam7."""
    import struct, zlib
    h, w, _ = rgb.shape
    if palette:
        colours, idx = np.unique(rgb.reshape(-1, 3), axis=0, return_inverse=True)
        assert len(colours) <= 256
        px = idx.reshape(h, w, 1).astype(np.uint8)
        ctype = 3
    elif alpha:
        px = np.concatenate([rgb, np.full((h, w, 1), 200, np.uint8)], axis=2)
        ctype = 6
    else:
        px, ctype = rgb, 2
    bpp = px.shape[2]
    raw = bytearray()
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    for xs, ys, dx, dy in passes:
        sub = px[ys::dy, xs::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        _png_filter_rows(sub, bpp, filter_types, raw)
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    z = zlib.compress(bytes(raw), 6)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 1 if interlace else 0))
    if palette:
        out += chunk(b"PLTE", colours.astype(np.uint8).tobytes())
    out += chunk(b"IDAT", z[: len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    return out


def _png_packed(idx, depth, ctype, interlace, palette=None):
    """Grey (ctype 0) or palette (ctype 3) PNG of `depth` bits per sample, samples packed MSB first, filter 0 / 1 / 2 by row."""
    import struct, zlib
    h, w = idx.shape
    raw = bytearray()
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    for xs, ys, dx, dy in passes:
        sub = idx[ys::dy, xs::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        ph, pw = sub.shape
        nbytes = (pw * depth + 7) // 8
        bits = np.zeros((ph, nbytes * 8), np.uint8)
        for b in range(depth):
            bits[:, b:pw * depth:depth] = (sub >> (depth - 1 - b)) & 1
        rows = np.packbits(bits, axis=1).astype(np.int32)
        prev = np.zeros(nbytes, np.int32)
        for y in range(ph):
            cur = rows[y]
            ft = y % 3
            pred = 0 if ft == 0 else (np.concatenate([[0], cur[:-1]]) if ft == 1 else prev)
            raw.append(ft)
            raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
            prev = cur
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    return out + chunk(b"IDAT", zlib.compress(bytes(raw), 6)) + chunk(b"IEND", b"")


def _png_wide(rgb16):
    """RGB PNG of 16 bits per sample (big-endian), filters 0 / 1 / 2 by row."""
    import struct, zlib
    h, w, _ = rgb16.shape
    rows = rgb16.astype(">u2").reshape(h, w * 3).view(np.uint8).reshape(h, w * 6).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * 6, np.int32)
    for y in range(h):
        cur = rows[y]
        ft = y % 3
        pred = 0 if ft == 0 else (np.concatenate([np.zeros(6, np.int32), cur[:-6]]) if ft == 1 else prev)
        raw.append(ft)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw), 6)) +
            chunk(b"IEND", b""))


def _png_filter_rows(px, bpp, filter_types, raw):
    h, w = px.shape[0], px.shape[1]
    rows = np.ascontiguousarray(px).reshape(h, w * bpp).astype(np.int32)
    prev = np.zeros(w * bpp, np.int32)
    for y in range(h):
        cur = rows[y]
        ft = filter_types[y % len(filter_types)]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        b = prev
        if ft == 0: pred = 0
        elif ft == 1: pred = a
        elif ft == 2: pred = b
        elif ft == 3: pred = (a + b) >> 1
        else:
            p = a + b - c
            pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        raw.append(ft)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur


def test_texture_file_decoders(ofdg, tmp_path):
    """PPM / BMP / PNG files decode to the B,G,R planes the reference holds after its channel swap (DataGenerator.cpp:128-133)."""
    import struct
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    rgb[5:20, 7:30] = rgb[5, 7]  # flat area: the PNG predictors see zeros
    want = np.ascontiguousarray(rgb[:, :, ::-1].transpose(2, 0, 1))
    (tmp_path / "a.ppm").write_bytes(b"P6\n# c\n53 37\n255\n" + rgb.tobytes())
    stride = (53 * 3 + 3) // 4 * 4
    rows = b"".join(rgb[y, :, ::-1].tobytes() + b"\0" * (stride - 53 * 3) for y in range(36, -1, -1))
    (tmp_path / "a.bmp").write_bytes(b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) +
                                     struct.pack("<IiiHHIIiiII", 40, 53, 37, 1, 24, 0, len(rows), 2835, 2835, 0, 0) + rows)
    (tmp_path / "a.png").write_bytes(_png_bytes(rgb))
    (tmp_path / "b.png").write_bytes(_png_bytes(rgb, alpha=True))
    few = (rgb // 64) * 64
    (tmp_path / "c.png").write_bytes(_png_bytes(few, palette=True))
    for depth in (1, 2, 4, 8):                                                       # grey / palette images below 8 bits per sample
        for interlace in (False, True):
            g = rng.integers(0, 1 << depth, (37, 53), dtype=np.uint8)
            (tmp_path / "g.png").write_bytes(_png_packed(g, depth, 0, interlace))
            grey = (g.astype(np.int32) * (255 // ((1 << depth) - 1))).astype(np.uint8)
            assert np.array_equal(ofdg.decode_texture_file(tmp_path / "g.png"), np.stack([grey] * 3)), (depth, interlace)
            pal = rng.integers(0, 256, (1 << depth, 3), dtype=np.uint8)
            (tmp_path / "p.png").write_bytes(_png_packed(g, depth, 3, interlace, pal))
            assert np.array_equal(ofdg.decode_texture_file(tmp_path / "p.png"), np.ascontiguousarray(pal[g][:, :, ::-1].transpose(2, 0, 1))), (depth, interlace)
    wide = rng.integers(0, 65536, (37, 53, 3), dtype=np.uint16)                      # 16 bits per sample: CImg<unsigned char> keeps the low byte
    (tmp_path / "w.png").write_bytes(_png_wide(wide))
    assert np.array_equal(ofdg.decode_texture_file(tmp_path / "w.png"), np.ascontiguousarray((wide & 255).astype(np.uint8)[:, :, ::-1].transpose(2, 0, 1)))
    (tmp_path / "d.png").write_bytes(_png_bytes(rgb, interlace=True))               # Adam7 (53 x 37: ragged passes)
    (tmp_path / "e.png").write_bytes(_png_bytes(rgb[:3, :2], interlace=True))       # passes without pixels
    assert np.array_equal(ofdg.decode_texture_file(tmp_path / "d.png"), want)
    assert np.array_equal(ofdg.decode_texture_file(tmp_path / "e.png"), np.ascontiguousarray(rgb[:3, :2, ::-1].transpose(2, 0, 1)))
    for name in ("a.ppm", "a.bmp", "a.png", "b.png"):
        assert np.array_equal(ofdg.decode_texture_file(tmp_path / name), want), name
    assert np.array_equal(ofdg.decode_texture_file(tmp_path / "c.png"), np.ascontiguousarray(few[:, :, ::-1].transpose(2, 0, 1)))
    (tmp_path / "x.gif").write_bytes(b"GIF89a" + b"\0" * 64)
    with pytest.raises(ofdg.OfdgError, match="unsupported image format"):
        ofdg.decode_texture_file(tmp_path / "x.gif")
    (tmp_path / "x.jpg").write_bytes(b"\xff\xd8\xff\xe0" + b"\0" * 64)  # a JPEG signature goes to nvJPEG (needs a device; garbage is rejected)
    with pytest.raises(ofdg.OfdgError, match="nvjpeg|JPEG"):
        ofdg.decode_texture_file(tmp_path / "x.jpg")
    with pytest.raises(ofdg.OfdgError, match="Could not open"):
        ofdg.decode_texture_file(tmp_path / "missing.png")
    (tmp_path / "t.ppm").write_bytes(b"P6\n53 37\n255\n" + rgb.tobytes()[:100])
    with pytest.raises(ofdg.OfdgError, match="truncated"):
        ofdg.decode_texture_file(tmp_path / "t.ppm")


@pytest.mark.gpu
def test_layer_mode9_uses_a_generated_field_pool(ofdg, oracle):
    """Mode 9 through the layer: the GPU field producer fills a pool of 40 crops at set-up (the reference runs a CropGenerator,
    DataGenerator.cpp:1016-1019); the batches equal the oracle's render with the same pool and parameter stream."""
    proto = ('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 6 prefetch: 2 } '
             'data_generation_param { mode: 9 texture_dbases: "synthetic:8:1" } }')
    layer = ofdg.DataGenerationLayer(proto)
    layer.LayerSetUp()
    layer.Forward_gpu()
    got = [layer.top_cpu(i) for i in range(3)]
    g = ofdg.Generator(device=0, mode=9, max_batch=6)
    fields = g.generate_fields(1, 40)          # the layer's seed for seed = 0, rank 0
    g.close()
    tasks = ofdg.ParamStream(9, n_fields=40).generate(6)
    bp = tasks.arrays()["blueprints"]
    assert ((bp["do_warpfield_deformation"] != 0) & (bp["field_id"] >= 0)).any()
    ref = oracle.render(tasks.struct(), ofdg.synth_textures(8, 1024, 768, seed=1), mode=9, fields=fields)
    ok = np.isfinite(ref["flow"])
    assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1
    assert np.array_equal(ok, np.isfinite(got[2])) and np.abs(got[2][ok] - ref["flow"][ok]).max() <= 1e-3
    layer.close()


@pytest.mark.gpu
def test_layer_mode9_refreshes_its_field_pool(ofdg, oracle):
    """The reference's CropGenerator keeps producing crops while training runs and hands every crop out three times
    (WarpFields.cpp:516-538, 540-641). The layer's pool is a ring of generations of 40 crops that the producer thread
    regenerates on the GPU: after enough batches the picks wrap around the ring and must find NEW crops in the old slots.
    Checked against the oracle's render with the crops the ring arithmetic says are resident (seed of generation G known)."""
    B, prefetch = 6, 2
    proto = ('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: %d prefetch: %d } '
             'data_generation_param { mode: 9 texture_dbases: "synthetic:8:1" } }' % (B, prefetch))
    ring = int(min(96.0, np.ceil(B * (23 * 0.2 + 0.2) / 120.0 * (prefetch + 3.0)) + 4.0))  # csrc/host/layer.cpp
    layer = ofdg.DataGenerationLayer(proto)
    layer.LayerSetUp()
    ps = ofdg.ParamStream(9, n_fields=40 * ring)
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    g = ofdg.Generator(device=0, mode=9, max_batch=B)
    gen_cache = {}

    def generation(G):  # FieldSeed(G) for seed 0, rank 0
        if G not in gen_cache:
            gen_cache[G] = g.generate_fields((1 + 104729 * G) & 0xFFFFFFFF, 40)
        return gen_cache[G]

    checked_wrapped = 0
    for it in range(60):
        d0 = ps.field_draws()
        tasks = ps.generate(B)
        d1 = ps.field_draws()
        layer.Forward_gpu()
        g_lo, g_hi = (d0 // 3) // 40, ((max(d1, d0 + 1) - 1) // 3) // 40
        if g_hi < ring or d1 == d0:
            continue  # still in the first lap of the ring (test_layer_mode9_uses_a_generated_field_pool covers it)
        got = [layer.top_cpu(i) for i in range(3)]
        fields = np.zeros((40 * ring, 2, 2, 385, 513), np.float32)
        for G in range(g_lo, g_hi + 1):
            fields[(G % ring) * 40:(G % ring) * 40 + 40] = generation(G)
        ref = oracle.render(tasks.struct(), tex, mode=9, fields=fields)
        ok = np.isfinite(ref["flow"])
        assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1, f"batch {it}"
        assert np.array_equal(ok, np.isfinite(got[2])) and np.abs(got[2][ok] - ref["flow"][ok]).max() <= 1e-3, f"batch {it}"
        checked_wrapped += 1
        if checked_wrapped == 2:
            break
    assert checked_wrapped == 2, "the picks never wrapped around the ring"
    g.close()
    layer.close()


@pytest.mark.gpu
def test_layer_forward_gpu_does_not_block_and_recycles_scenes(ofdg, oracle):
    """Forward_gpu queues the batch on the layer's stream and returns; the default stream (top_cpu's copy) is ordered after
    the blobs by an event. Many forwards back to back -- prepared scenes go back to the generator's free list and are reused --
    still deliver the commission order of the parameter stream."""
    import time
    text = ('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 8 prefetch: 3 } '
            'data_generation_param { mode: 7 texture_dbases: "synthetic:8:1" } }')
    layer = ofdg.DataGenerationLayer(text)
    layer.LayerSetUp()
    ps = ofdg.ParamStream(7)
    tex = ofdg.synth_textures(8, 1024, 768, seed=1)
    n = 40
    for _ in range(n - 1):
        layer.Forward_gpu()  # (no read in between: the blobs are simply overwritten in stream order)
        ps.generate(8)
    layer.Forward_gpu()
    got = [layer.top_cpu(i) for i in range(3)]
    ref = oracle.render(ps.generate(8).struct(), tex, mode=7)
    assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1
    assert np.abs(got[2] - ref["flow"]).max() <= 1e-3
    layer.close()


def test_texture_list_follows_the_reference_loop(ofdg, tmp_path):
    """TextureCollection's loop (DataGenerator.cpp:123-126) is `while (!eof) { getline; if (eof) break; load; }`: a last line
    without a trailing newline is NOT loaded, so the pool size -- and with it tex_id % pool size -- must follow the same rule."""
    f = tmp_path / "db.txt"
    f.write_text("a.ppm\nb.ppm\nc.ppm\n")
    assert ofdg.read_texture_list(f) == ["a.ppm", "b.ppm", "c.ppm"]
    f.write_text("a.ppm\nb.ppm\nc.ppm")           # unterminated last line: dropped, like the reference
    assert ofdg.read_texture_list(f) == ["a.ppm", "b.ppm"]
    f.write_text("a.ppm\r\nb.ppm\r\n")
    assert ofdg.read_texture_list(f) == ["a.ppm", "b.ppm"]
    f.write_text("a.ppm\n\nb.ppm\n")              # CImg::load("") throws in the reference
    with pytest.raises(ofdg.OfdgError, match="empty line"):
        ofdg.read_texture_list(f)
    f.write_text("only-line-without-newline")
    with pytest.raises(ofdg.OfdgError, match="empty"):
        ofdg.read_texture_list(f)
    with pytest.raises(ofdg.OfdgError, match="Could not open texture collection"):
        ofdg.read_texture_list(tmp_path / "missing.txt")


@pytest.mark.gpu
def test_jpeg_textures_through_nvjpeg(ofdg, tmp_path):
    """JPEG textures (the authors' database format; CImg::load -> libjpeg in the reference, DataGenerator.cpp:128) are decoded by
    nvJPEG. Checked against Pillow's libjpeg decode of the same files: baseline 4:2:0, progressive 4:4:4 and grayscale."""
    from PIL import Image
    tex = ofdg.synth_textures(1, 640, 480, seed=21)[0]               # planes in file order R, G, B
    rgb = np.ascontiguousarray(tex.transpose(1, 2, 0))
    Image.fromarray(rgb).save(tmp_path / "a.jpg", quality=92, subsampling=2)
    Image.fromarray(rgb).save(tmp_path / "b.jpg", quality=95, subsampling=0, progressive=True)
    Image.fromarray(rgb[:, :, 1]).save(tmp_path / "c.jpg", quality=90)
    for name in ("a.jpg", "b.jpg", "c.jpg"):
        planar = ofdg.decode_texture_file(tmp_path / name)
        want = np.asarray(Image.open(tmp_path / name).convert("RGB")).astype(int)
        assert planar.shape == (3, 480, 640)
        got = planar[::-1].transpose(1, 2, 0).astype(int)            # planar B,G,R -> interleaved R,G,B
        d = np.abs(got - want)
        # decoders differ in IDCT rounding and, for subsampled chroma (a.jpg, 4:2:0), in the upsampling filter: libjpeg's "fancy"
        # triangle filter against nvJPEG's (measured: max 6 / mean 0.77 on 4:2:0, max 4 / mean 0.52 on progressive 4:4:4)
        lim_max, lim_mean = (16, 1.5) if name == "a.jpg" else (8, 1.0)
        assert d.max() <= lim_max and d.mean() < lim_mean, (name, d.max(), d.mean())
        if name != "c.jpg":
            assert np.abs(got - rgb.astype(int)).mean() < 8  # and it is the picture that was encoded
    # and through the layer's list loader
    (tmp_path / "db.txt").write_text(str(tmp_path / "a.jpg") + "\n" + str(tmp_path / "b.jpg") + "\n")
    proto = ('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 2 prefetch: 1 } '
             'data_generation_param { mode: 5 texture_dbases: "%s" } }' % (tmp_path / "db.txt"))
    layer = ofdg.DataGenerationLayer(proto)
    layer.LayerSetUp()
    layer.Forward_gpu()
    assert layer.top_cpu(0).std() > 5
    layer.close()
