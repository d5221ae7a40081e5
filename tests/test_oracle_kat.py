"""Known-answer and property tests that pin the CPU oracle (the reference ships no golden vectors:
SURVEY 4 / 8c). Hand-derived from the AGG 2.4 scanline algorithm."""
import numpy as np
import pytest


def test_rect_half_pixel_kat(oracle):
    # rect [0.5,2.5]x[0,1]: covers 128,255,128 -> gray8 LUT -> 127,255,127; threshold gamma -> 255 x3
    xy = [[0.5, 0], [2.5, 0], [2.5, 1], [0.5, 1]]
    aa = oracle.raster_polygon(xy, 8, 4, True)
    na = oracle.raster_polygon(xy, 8, 4, False)
    assert aa[0, :4].tolist() == [127, 255, 127, 0] and not aa[1:].any()
    assert na[0, :4].tolist() == [255, 255, 255, 0] and not na[1:].any()


def test_rect_integer_coords(oracle):
    # integer corners: exact 255 interior, nothing on the right/bottom edge cells
    m = oracle.raster_polygon([[1, 1], [4, 1], [4, 3], [1, 3]], 8, 4, True)
    ref = np.zeros((4, 8), np.uint8)
    ref[1:3, 1:4] = 255
    assert np.array_equal(m, ref)


def test_partial_cover_never_254(oracle):
    # graylut: a partially covered AA pixel is at most 253
    rng = np.random.default_rng(0)
    for _ in range(20):
        c = rng.uniform(10, 50, 2)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 7))
        r = rng.uniform(3, 9, 7)
        xy = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
        m = oracle.raster_polygon(xy, 64, 64, True)
        assert not (m == 254).any()
        assert (m == 255).any()


def _supersampled_cover(xy, W, H, ss=16):
    # even-odd == non-zero for simple polygons; sample centres of an ss x ss grid per pixel
    ys, xs = np.meshgrid((np.arange(H * ss) + 0.5) / ss, (np.arange(W * ss) + 0.5) / ss, indexing="ij")
    inside = np.zeros(xs.shape, bool)
    n = len(xy)
    for i in range(n):
        x1, y1 = xy[i]
        x2, y2 = xy[(i + 1) % n]
        if y1 == y2:
            continue
        cond = ((y1 <= ys) & (ys < y2)) | ((y2 <= ys) & (ys < y1))
        xint = x1 + (ys - y1) * (x2 - x1) / (y2 - y1)
        inside ^= cond & (xs < xint)
    return inside.reshape(H, ss, W, ss).mean((1, 3))


@pytest.mark.parametrize("seed", range(6))
def test_coverage_matches_geometric_area(oracle, seed):
    # AGG's cell accumulation is the exact covered area per pixel up to 1/256 rounding
    rng = np.random.default_rng(seed)
    W, H = 48, 40
    c = rng.uniform([5, 5], [W - 5, H - 5])
    k = int(rng.integers(3, 9))
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    if seed % 2:
        ang = ang[::-1]  # both windings
    r = rng.uniform(4, 22, k)  # may leave the frame on every side (clipping left of x=0 included)
    xy = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
    m = oracle.raster_polygon(xy, W, H, True).astype(np.float64) / 255
    ref = _supersampled_cover(xy, W, H)
    assert np.abs(m - ref).max() < 0.08  # supersampling error dominates
    assert abs(m.sum() - ref.sum()) / max(ref.sum(), 1) < 0.01
    na = oracle.raster_polygon(xy, W, H, False)
    assert set(np.unique(na)) <= {0, 255}
    assert np.array_equal(na == 255, oracle.raster_polygon(xy, W, H, True) >= 127)


def test_winding_changes_at_most_one_level(oracle):
    # the arithmetic >>9 before abs() makes CW and CCW differ by rounding only
    xy = np.array([[3.3, 2.1], [20.7, 5.9], [11.2, 17.4]])
    a = oracle.raster_polygon(xy, 32, 24, True).astype(int)
    b = oracle.raster_polygon(xy[::-1], 32, 24, True).astype(int)
    assert np.abs(a - b).max() <= 2
    assert (a != b).any()  # and it is visible somewhere, which is why the sign has to be kept


def test_fixed_point_entry_matches_double_entry(oracle):
    xy = np.array([[3.3, 2.1], [20.7, 5.9], [11.2, 17.4]])
    fx = np.where(xy * 256 < 0, xy * 256 - 0.5, xy * 256 + 0.5).astype(np.int32)  # agg::iround
    assert np.array_equal(oracle.raster_polygon(xy, 32, 24, True), oracle.raster_fixed(fx, 32, 24, True))


def test_composite_luts(oracle):
    add, sub = oracle.composite_luts()
    assert add[0, 128] == 128 and sub[255, 128] == 126
    for u in (0, 255):
        for v in (0, 255):
            assert add[u, v] in (0, 255) and sub[u, v] in (0, 255)
    assert add[0, 0] == 0 and sub[0, 77] == 0
    # the float formula is NOT the identity for v == 0: the GPU must apply it to untouched pixels too
    assert (add[:, 0] != np.arange(256)).any()


def test_draw_image_float_equals_integer_floor():
    # CImg draw_image: (T)((|m|*s + d*(255-m))/255) in float == floor((m*s + d*(255-m))/255)  
    m = np.arange(256, dtype=np.float32)[:, None, None]
    s = np.arange(256, dtype=np.float32)[None, :, None]
    d = np.arange(256, dtype=np.float32)[None, None, :]
    f = ((m * s + d * (np.float32(255) - m)) / np.float32(255)).astype(np.uint8)
    mi, si, di = (a.astype(np.int64) for a in (m, s, d))
    assert np.array_equal(f, ((mi * si + di * (255 - mi)) // 255).astype(np.uint8))


def test_dda2_closed_form():
    # dda2_line_interpolator stepped i times == v1 + i*lft + ceil((i+1)*rem/n) - 1  
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.choice([512, 1024, 7, 100]))
        v1, v2 = (int(v) for v in rng.integers(-400000, 400000, 2))
        d = v2 - v1
        lft, rem = int(d / n), int(np.fmod(d, n))  # C truncation
        mod = rem
        if mod <= 0:
            mod += n; rem += n; lft -= 1
        mod -= n
        y, seq = v1, []
        for i in range(n):
            seq.append(y)
            mod += rem; y += lft
            if mod > 0:
                mod -= n; y += 1
        closed = [v1 + i * lft + ((i + 1) * rem + n - 1) // n - 1 for i in range(n)]
        assert seq == closed


def test_identity_warp_is_a_copy(oracle, textures8):
    img = textures8[0][:, :96, :128].copy()
    assert np.array_equal(oracle.transform_texture(img, [1, 0, 0, 1, 0, 0]), img)


def test_integer_translation_warp_reflects(oracle, textures8):
    img = textures8[1][:, :64, :80].copy()
    out = oracle.transform_texture(img, [1, 0, 0, 1, 5, -3])  # dst(x,y) = src(x-5, y+3), reflect outside
    xs = np.arange(80) - 5
    ys = np.arange(64) + 3
    rx = np.where(xs < 0, -xs - 1, xs)
    ry = np.where(ys >= 64, 2 * 64 - ys - 1, ys)
    assert np.array_equal(out, img[:, ry][:, :, rx])


def test_default_randomized_crop_is_centre_crop(oracle, textures8):
    t = textures8[2]
    c = oracle.randomized_crop(t, 512, 384)
    assert np.array_equal(c, t[:, 192:576, 256:768])


def test_background_crop_paths(oracle, textures8):
    t = textures8[3]
    a = oracle.randomized_crop(t, 1024, 768, 0.0, 1.0, 0, 0)
    assert np.array_equal(a, t)  # no shift, no rotation, zoom 1: the chain is the identity
    b = oracle.randomized_crop(t, 1024, 768, 0.0, 1.0, 512, 0)
    assert np.array_equal(b[:, :, 512:], t[:, :, :512]) and np.array_equal(b[:, :, :512], t[:, :, 511::-1])
    for zoom in (0.83, 1.17):  # moving average / linear
        c = oracle.randomized_crop(t, 1024, 768, 2.5, zoom, 0, 384)
        assert c.shape == (3, 768, 1024) and c.std() > 5
