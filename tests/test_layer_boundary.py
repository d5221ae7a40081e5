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
    with pytest.raises(ofdg.OfdgError, match="DataGeneration"):
        ofdg.DataGenerationLayer('layer { type: "Data" top: "a" }')


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
def test_layer_ppm_texture_list(ofdg, oracle, tmp_path):
    """texture_dbases as a list file of PPM images, with the reference's R<->B swap."""
    rng = np.random.default_rng(0)
    tex_rgb = ofdg.synth_textures(2, 1024, 768, seed=9)  # treat planes as R,G,B on disk
    lines = []
    for i in range(2):
        path = tmp_path / f"t{i}.ppm"
        with open(path, "wb") as f:
            f.write(b"P6\n# comment\n1024 768\n255\n")
            f.write(np.ascontiguousarray(tex_rgb[i].transpose(1, 2, 0)).tobytes())
        lines.append(str(path))
    (tmp_path / "db.txt").write_text("\n".join(lines) + "\n")
    layer = ofdg.DataGenerationLayer('layer { type: "DataGeneration" top: "a" top: "b" top: "c" data_param { batch_size: 2 prefetch: 2 } '
                                     'data_generation_param { mode: 5 texture_dbases: "%s" } }' % (tmp_path / "db.txt"))
    layer.LayerSetUp()
    layer.Forward_gpu()
    got = [layer.top_cpu(i) for i in range(3)]
    tasks = ofdg.ParamStream(5).generate(2)
    ref = oracle.render(tasks.struct(), tex_rgb[:, ::-1].copy(), mode=5)  # planes held as B,G,R
    assert np.abs(got[0] - ref["img0"]).max() <= 1 and np.abs(got[1] - ref["img1"]).max() <= 1
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
