"""The pair binning and the mask rasterisation of a batch run on a high-priority side stream beside the background
preparation (csrc/api.cu:run_kernels), and consecutive batches are software-pipelined over two scratch sets: the front end
(preparation, binning, rasterisation) of batch k+1 runs beside the shade kernel of batch k (run_kernels_pipelined). Where and when
a kernel runs must not change a single bit of the blobs: batches queued back to back on a caller's stream without any host
synchronisation in between (the front end of batch k+2 must wait for the shade kernel of batch k, which still reads the scratch
set they share) are compared with the same batches rendered with everything in line on one stream
(OFDG_BIN_OVERLAP=0 OFDG_PIPELINE=0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_BATCHES, BATCH = 8, 16


def _blobs(n):
    import torch
    return (torch.empty((n, 3, 384, 512), device="cuda"), torch.empty((n, 3, 384, 512), device="cuda"),
            torch.empty((n, 2, 384, 512), device="cuda"))


def _run(ofdg, textures8, mode, philox, mixed=False):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    g = ofdg.Generator(device=0, mode=mode, max_batch=BATCH)
    g.upload_textures(textures8)
    ts = torch.cuda.Stream()
    outs = [_blobs(BATCH) for _ in range(N_BATCHES)]
    if philox:
        for k, (i0, i1, fl) in enumerate(outs):
            g.generate_philox(77, k * BATCH, BATCH, i0, i1, fl, stream=ts.cuda_stream)
    else:
        ps = ofdg.ParamStream(mode)
        prepared = [g.prepare(ps.generate(BATCH)) for _ in range(N_BATCHES)]
        ts2 = torch.cuda.Stream()
        ps2 = ofdg.ParamStream(mode)
        for k, (p, (i0, i1, fl)) in enumerate(zip(prepared, outs)):
            if mixed and k % 3 == 1:  # an in-order call (fresh scene upload, scratch set 0) between pipelined ones
                ps2.skip(k * BATCH - ps2.tasks_generated())
                g.render(ps2.generate(BATCH), i0, i1, fl, ts.cuda_stream)
            elif mixed and k % 3 == 2:  # a pipelined call on another caller stream
                g.render_prepared(p, i0, i1, fl, ts2.cuda_stream)
            else:
                g.render_prepared(p, i0, i1, fl, ts.cuda_stream)
        ts2.synchronize()
    ts.synchronize()
    torch.cuda.synchronize()
    host = [[t.cpu().numpy() for t in o] for o in outs]
    g.close()
    return host


@pytest.mark.parametrize("philox", [False, True])
@pytest.mark.parametrize("mode", [7, 2])
def test_side_stream_changes_no_bit(ofdg, textures8, mode, philox, monkeypatch):
    monkeypatch.setenv("OFDG_BIN_OVERLAP", "0")
    monkeypatch.setenv("OFDG_PIPELINE", "0")
    inline = _run(ofdg, textures8, mode, philox)
    monkeypatch.delenv("OFDG_BIN_OVERLAP")
    monkeypatch.delenv("OFDG_PIPELINE")
    monkeypatch.setenv("OFDG_PHILOX_PIPELINE", "1")  # (off by default for the device stream; its pipelined form stays covered)
    forked = _run(ofdg, textures8, mode, philox)
    assert inline[0][0].std() > 10
    assert not np.array_equal(inline[0][2], inline[1][2]), "the batches are supposed to differ"
    for k, (a, b) in enumerate(zip(inline, forked)):
        for name, x, y in zip(("img0", "img1", "flow"), a, b):
            assert np.array_equal(x, y), f"batch {k}: {name} differs between the in-line and the forked step"


def test_pipelined_and_in_order_calls_mix(ofdg, textures8, monkeypatch):
    """Pipelined calls (prepared scenes, alternating scratch sets), in-order calls (ofdg_render: scratch set 0) and a second
    caller stream interleaved without host synchronisation: every batch equals its in-line rendering."""
    monkeypatch.setenv("OFDG_BIN_OVERLAP", "0")
    monkeypatch.setenv("OFDG_PIPELINE", "0")
    inline = _run(ofdg, textures8, 7, False)
    monkeypatch.delenv("OFDG_BIN_OVERLAP")
    monkeypatch.delenv("OFDG_PIPELINE")
    mixed = _run(ofdg, textures8, 7, False, mixed=True)
    for k, (a, b) in enumerate(zip(inline, mixed)):
        for name, x, y in zip(("img0", "img1", "flow"), a, b):
            assert np.array_equal(x, y), f"batch {k}: {name} differs between the in-line and the mixed pipelined step"
