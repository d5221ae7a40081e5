"""The render kernel's closed-form tile rasteriser (csrc/raster_tile.h), executed on the host through
the C ABI, against the oracle's sequential AGG port: bit-exact on every pixel, AA and non-AA."""
import numpy as np
import pytest


def _fixed(xy):
    xy = np.asarray(xy, np.float64) * 256
    return np.where(xy < 0, xy - 0.5, xy + 0.5).astype(np.int32)


def _check(ofdg, oracle, fx, W, H):
    for aa in (True, False):
        ref = oracle.raster_fixed(fx, W, H, aa)
        got = ofdg.raster_host(fx, W, H, aa)
        assert np.array_equal(ref, got), f"aa={aa}: {np.argwhere(ref != got)[:5]}"


@pytest.mark.parametrize("seed", range(40))
def test_random_polygons(ofdg, oracle, seed):
    rng = np.random.default_rng(seed)
    W, H = 512, 40
    k = int(rng.integers(3, 24))
    # star-ish and self-intersecting outlines, any winding, partly (or wholly) outside the frame
    c = rng.uniform([-200, -20], [700, 60])
    ang = rng.uniform(0, 2 * np.pi, k)
    if seed % 3:
        ang = np.sort(ang)
    r = rng.uniform(1, 400 if seed % 2 else 60, k)
    xy = np.stack([c[0] + r * np.cos(ang), c[1] + 0.2 * r * np.sin(ang)], 1)
    _check(ofdg, oracle, _fixed(xy), W, H)


def test_axis_aligned_and_degenerate_edges(ofdg, oracle):
    W, H = 300, 24
    cases = [
        [[0.5, 0], [2.5, 0], [2.5, 1], [0.5, 1]],
        [[1, 1], [260, 1], [260, 20], [1, 20]],                    # crosses tile columns 0..2, integer corners
        [[-50.3, -10.2], [350.7, -10.2], [350.7, 40.9], [-50.3, 40.9]],  # covers the whole frame
        [[10, 5], [10, 5], [200.25, 5.5], [200.25, 5.5], [100, 18.75]],  # duplicate vertices
        [[127.99, 3.01], [128.01, 3.01], [128.01, 12.99], [127.99, 12.99]],  # sliver across a tile boundary
        [[5, 7.5], [290, 7.6], [290, 7.7], [5, 7.6]],              # nearly horizontal: hundreds of cells per row segment
        [[-1000, 3], [-900, 3], [-900, 20], [-1000, 20]],          # entirely left of the frame
        [[400, 3], [500, 3], [500, 20], [400, 20]],                # entirely right of the frame
        [[-300.5, 2.25], [250.5, 21.75], [-300.5, 21.75]],          # starts far left of x = 0 (clipping is off)
    ]
    for xy in cases:
        _check(ofdg, oracle, _fixed(xy), W, H)
        _check(ofdg, oracle, _fixed(xy[::-1]), W, H)


def test_flattened_shapes(ofdg, oracle):
    """Ellipses and curved polygons as the host geometry stage emits them."""
    rng = np.random.default_rng(7)
    W, H = 512, 384
    for i in range(12):
        a = rng.uniform(-np.pi, np.pi)
        m = [np.cos(a), np.sin(a), -np.sin(a), np.cos(a), rng.uniform(-100, 600), rng.uniform(-100, 480)]
        if i % 2:
            fx = ofdg.flatten_ellipse(rng.uniform(2, 100), rng.uniform(2, 100), m)
            assert len(fx) == 100
        else:
            n = int(rng.integers(3, 21))
            phi = np.arange(n) * 2 * np.pi / n
            rad = rng.uniform(20, 160, n)
            st = [ofdg.SEG_DUMMY] + [ofdg.SEG_LINE] * (n - 1)
            j = 1
            while j < n - 1:
                if rng.uniform() < 0.4:
                    st[j], st[j + 1] = ofdg.SEG_CURVE3, ofdg.SEG_DUMMY
                    j += 2
                else:
                    j += 1
            fx = ofdg.flatten_polygon(st, rad * np.cos(phi), rad * np.sin(phi), m)
            assert len(fx) >= n - st.count(ofdg.SEG_CURVE3)
        _check(ofdg, oracle, fx, W, H)
