// sm_100a kernels of the generator's render path. Build with --fmad=false: every float/double
// expression below is meant to round exactly like the reference's x86-64 (SSE, no FMA) build.
//
// What each kernel replaces in /root/reference/src/caffe/DataGenerator.cpp ("DG.cpp"):
//   bg_tables_kernel / bg_prep_kernel  Texture::getRandomizedCrop(2W,2H,..) for the background
//                                      (DG.cpp:87-109, 1186-1192; CImg shift/rotate/crop/resize)
//   render_kernel                      per 128x8-pixel tile, everything else of Process_TaskBucket:
//       raster   MovingObjectBase::draw<> = AGG scanline cell/cover accumulation  (DG.cpp:351-368)
//       combine  MovingObjectComposite::renderMasks                                  (DG.cpp:591-646)
//       warp     getTransformedTexture = AGG span bilinear filter, reflect wrap      (DG.cpp:168-231)
//       blit     RenderCore::blitObject = object ids + CImg draw_image blend         (DG.cpp:762-799)
//       flow     RenderCore::computeFlowImage / getPointFlow                          (DG.cpp:801-818, 388-407, 692-718)
//       write    uint8 -> float conversion into the output blobs                      (DG.cpp:1229-1245)
#include "render.cuh"

#include <cstdlib>
#include <string>

#include "ofdg/augment.h"
#include "raster_tile.h"

namespace ofdg {

namespace {

constexpr int RENDER_THREADS = 32 * TH;  // TH warps; lane = 4 consecutive pixels (one 128-bit store per lane and plane)
#ifndef OFDG_RENDER_MIN_BLOCKS
#define OFDG_RENDER_MIN_BLOCKS 3
#endif

// ------------------------------------------------------------------------------------------------
// small integer helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int iround_d(double v) { return int((v < 0.0) ? v - 0.5 : v + 0.5); }

// agg::wrap_mode_reflect. The generic modulo form is kept out of line: every tap near the image needs one fold at most,
// and the render kernel must stay small enough for the instruction cache.
__device__ __noinline__ int reflect_far(int v, int size) {
  unsigned size2 = 2u * (unsigned)size;
  unsigned add = size2 * (0x3FFFFFFFu / size2);
  unsigned m = ((unsigned)v + add) % size2;
  return (int)(m >= (unsigned)size ? size2 - m - 1 : m);
}
__device__ __forceinline__ int reflect(int v, int size) {
  if ((unsigned)v < (unsigned)size) return v;
  if (v < 0 && v >= -size) return -v - 1;
  if (v >= size && v < 2 * size) return 2 * size - v - 1;
  return reflect_far(v, size);
}
// CImg mirror boundary: cimg::mod(i, 2n), then fold
__device__ __forceinline__ int mirror(int i, int n) {
  if ((unsigned)i < (unsigned)n) return i;
  if (i < 0 && i >= -n) return -i - 1;
  if (i >= n && i < 2 * n) return 2 * n - i - 1;
  int n2 = 2 * n;
  int m = i % n2;
  if (m < 0) m += n2;
  return m < n ? m : n2 - m - 1;
}

// dda2_line_interpolator in closed form: value after i increments
struct Dda2 {
  int v1, lft, rem, n, sh;
  __device__ __forceinline__ void init(int a, int b, int count) {
    n = count;
    sh = (count & (count - 1)) == 0 ? 31 - __clz(count) : -1;  // 512 / 1024 wide frames: shift instead of divide
    int d = b - a;
    if (sh >= 0) {  // C division truncates toward zero
      lft = (d + ((d >> 31) & (n - 1))) >> sh;
      rem = d - (lft << sh);
    } else {
      lft = d / n;
      rem = d % n;
    }
    if (rem <= 0) { rem += n; lft--; }
    v1 = a;
  }
  __device__ __forceinline__ int at(int i) const {
    const int t = (i + 1) * rem + n - 1;  // > 0
    return v1 + i * lft + (sh >= 0 ? (t >> sh) : t / n) - 1;
  }
};

// One row of agg::span_interpolator_linear + span_image_filter_rgb_bilinear over an RGBX image
// addressed as img[(oy + r(y)) * pitch + ox + r(x)], r = reflect over (sw, sh).
struct RowWarp {
  Dda2 dx, dy;
  __device__ __forceinline__ void init(const double* m, double y, int n) {
    // begin(0.5, y + 0.5, n): endpoints through the inverse matrix, iround(256 * .)
    double tx = 0.5, ty = y + 0.5;
    double ax = tx * m[0] + ty * m[2] + m[4];
    double ay = tx * m[1] + ty * m[3] + m[5];
    int X1 = iround_d(ax * 256.0), Y1 = iround_d(ay * 256.0);
    tx = 0.5 + n;
    double bx = tx * m[0] + ty * m[2] + m[4];
    double by = tx * m[1] + ty * m[3] + m[5];
    int X2 = iround_d(bx * 256.0), Y2 = iround_d(by * 256.0);
    dx.init(X1, X2, n);
    dy.init(Y1, Y2, n);
  }
};

// byte c of a packed pixel as float without the (quarter-rate) I2F unit: 0x4B0000xx is 8388608 + xx exactly
__device__ __forceinline__ float byte_to_float(uint32_t px, int c) {
  return __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7440u + (unsigned)c)) - 8388608.0f;
}

// The same through the conversion unit (I2F.U8 with a byte selector: one instruction instead of two, on a pipe the shade
// kernel otherwise leaves idle -- it is latency-bound since round 2, not issue-bound)
__device__ __forceinline__ float byte_to_float_cvt(uint32_t px, int c) { return (float)((px >> (8 * c)) & 255u); }
#ifndef OFDG_SHADE_VEC_FG
#define OFDG_SHADE_VEC_FG 1  // the lane's four frame-0 texels of an object as one 128-bit load where the view is 16-byte aligned (0.1334 -> 0.1329 ms)
#endif
#ifndef OFDG_SHADE_I2F
#define OFDG_SHADE_I2F 2  // 0: permute + add for both frames, 1: frame 0 through the conversion unit, 2: both frames (0.1334 / 0.1332 / 0.1328 ms)
#endif
__device__ __forceinline__ float byte_to_biased(uint32_t px, int c) {  // 2^23 + byte c (exact)
  return __uint_as_float(__byte_perm(px, 0x4B000000u, 0x7440u + (unsigned)c));
}
__device__ __forceinline__ uint32_t ld_px(const uchar4* p) { return *reinterpret_cast<const uint32_t*>(p); }

__device__ __forceinline__ uint32_t bilinear_rgbx(const uchar4* img, int pitch, int ox, int oy, int sw, int sh,
                                                  const RowWarp& rw, int i) {
  int x_hr = rw.dx.at(i) - 128, y_hr = rw.dy.at(i) - 128;
  int x_lr = x_hr >> 8, y_lr = y_hr >> 8;
  unsigned fx = x_hr & 255, fy = y_hr & 255;
  int xa = reflect(x_lr, sw), xb = reflect(x_lr + 1, sw), ya = reflect(y_lr, sh), yb = reflect(y_lr + 1, sh);
  const uchar4* r0 = img + (size_t)(oy + ya) * pitch + ox;
  const uchar4* r1 = img + (size_t)(oy + yb) * pitch + ox;
  uint32_t p00 = ld_px(r0 + xa), p10 = ld_px(r0 + xb), p01 = ld_px(r1 + xa), p11 = ld_px(r1 + xb);
  // sum w_k p_k = [p00 (256-fx) + p10 fx] (256-fy) + [p01 (256-fx) + p11 fx] fy, exact in integers;
  // the horizontal pair is one dp4a: bytes {p00_c, p10_c, p00_c, 0} . {255-fx, fx, 1, 0}
  const unsigned wx = (255u - fx) | (fx << 8) | (1u << 16);
  unsigned sacc[3];  // below 2^24: byte 2 is the channel's result, byte 3 is zero
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned sel = (unsigned)c | ((4u + c) << 4) | ((unsigned)c << 8) | (3u << 12);
    const unsigned top = __dp4a(__byte_perm(p00, p10, sel), wx, 0u);
    const unsigned bot = __dp4a(__byte_perm(p01, p11, sel), wx, 0u);
    sacc[c] = 32768u + top * (256u - fy) + bot * fy;
  }
  return __byte_perm(__byte_perm(sacc[0], sacc[1], 0x0062u), sacc[2], 0x7610u);
}

// Out-of-line copy for the render kernel's eight call sites (keeps its code inside the instruction cache);
// scalar arguments only, so that everything travels in registers. base = img + oy * pitch + ox.
__device__ __noinline__ uint32_t bilinear_rgbx_call(const uchar4* base, int pitch, int sw, int sh, int n, int v1x, int lftx, int remx,
                                                    int v1y, int lfty, int remy, int i) {
  RowWarp rw;
  const int shift = (n & (n - 1)) == 0 ? 31 - __clz(n) : -1;
  rw.dx.v1 = v1x; rw.dx.lft = lftx; rw.dx.rem = remx; rw.dx.n = n; rw.dx.sh = shift;
  rw.dy.v1 = v1y; rw.dy.lft = lfty; rw.dy.rem = remy; rw.dy.n = n; rw.dy.sh = shift;
  return bilinear_rgbx(base, pitch, 0, 0, sw, sh, rw, i);
}

// The same filter for a lane's four consecutive span pixels when (a) the span length is a power of two, so the closed-form
// dda2 value (Dda2::at) advances by an add and a shift per pixel, and (b) every tap of every lane of the warp lies inside
// the image, so wrap_mode_reflect is the identity (checked by the caller: span positions are monotone along the row, the
// lane's first and last pixel bound the rest). The rows' interpolator constants come from a table (RenderArgs::bg_rows /
// pair_rows) instead of being worked out by every warp.
struct SpanLane {
  int xb, yb, tx, ty, lx, ly, rx, ry, sh;  // x_hr(k) = xb + k * lx + ((tx + k * rx) >> sh), k = 0..3 (the -128 of the filter offset is in xb)
  __device__ __forceinline__ void init(const int4 r0, const int4 r1, int i0, int n, int shift) {
    lx = r0.y; rx = r0.z; ly = r1.x; ry = r1.y; sh = shift;
    tx = (i0 + 1) * rx + n - 1; ty = (i0 + 1) * ry + n - 1;
    xb = r0.x + i0 * lx - 129; yb = r0.w + i0 * ly - 129;
  }
  __device__ __forceinline__ int x_hr(int k) const { return xb + k * lx + ((tx + k * rx) >> sh); }
  __device__ __forceinline__ int y_hr(int k) const { return yb + k * ly + ((ty + k * ry) >> sh); }
  // every tap of pixels 0..3 inside [0, sw - 1] x [0, sh_img - 1]?
  __device__ __forceinline__ bool inside(int sw, int sh_img) const {
    const int xa = x_hr(0) >> 8, xz = x_hr(3) >> 8, ya = y_hr(0) >> 8, yz = y_hr(3) >> 8;
    return (unsigned)xa < (unsigned)(sw - 1) && (unsigned)xz < (unsigned)(sw - 1) && (unsigned)ya < (unsigned)(sh_img - 1) && (unsigned)yz < (unsigned)(sh_img - 1);
  }
};
__device__ __forceinline__ uint32_t bilinear_inside(const uint32_t* img, int pitch, int x_hr, int y_hr) {
  const int x_lr = x_hr >> 8, y_lr = y_hr >> 8;
  const unsigned fx = x_hr & 255, fy = y_hr & 255;
  const uint32_t* r0 = img + (y_lr * pitch + x_lr);  // one texture (or one prepared background): the index fits 32 bits
  const uint32_t* r1 = r0 + pitch;
  const uint32_t p00 = r0[0], p10 = r0[1], p01 = r1[0], p11 = r1[1];
  const unsigned wx = (255u - fx) | (fx << 8) | (1u << 16);
  unsigned sacc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned sel = (unsigned)c | ((4u + c) << 4) | ((unsigned)c << 8) | (3u << 12);
    const unsigned top = __dp4a(__byte_perm(p00, p10, sel), wx, 0u);
    const unsigned bot = __dp4a(__byte_perm(p01, p11, sel), wx, 0u);
    sacc[c] = 32768u + top * (256u - fy) + bot * fy;
  }
  return __byte_perm(__byte_perm(sacc[0], sacc[1], 0x0062u), sacc[2], 0x7610u);
}
__device__ __forceinline__ int pow2_shift(int n) { return (n & (n - 1)) == 0 ? 31 - __clz(n) : -1; }

// CImg draw_image(sprite, mask, 1, 255) per channel == floor((m*t + f*(255-m)) / 255)  
__device__ __forceinline__ uint32_t blend_rgbx(uint32_t f, uint32_t t, unsigned m) {
  const unsigned w = m | ((255u - m) << 8);  // one dp4a per channel: bytes {t_c, f_c, 0, 0} . {m, 255 - m, 0, 0}
  unsigned q[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    q[c] = __umulhi(__dp4a(__byte_perm(t, f, (unsigned)c | ((4u + c) << 4)), w, 0u), 16843010u);  // floor(x / 255), x <= 65025
  return __byte_perm(__byte_perm(q[0], q[1], 0x0040u), q[2], 0x5410u);  // q <= 255: q[2]'s byte 1 supplies the zero
}

// The blob write of a lane's four pixels with the augmentation applied (both frames, three channels). Out of line: the
// reference path never takes it, and inlined its Philox rounds were most of the render kernels' static code. One Philox call
// per pixel and frame yields the three channels' noise sums (ofdg_noise3).
__device__ __noinline__ void store_augmented(const ofdg_augment* au, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                             uint32_t b2, uint32_t b3, float* o0, float* o1, size_t P, uint32_t p) {
  const uint32_t k0 = au->noise_seed[0], k1 = au->noise_seed[1];
  const uint32_t n00 = ofdg_noise3(k0, k1, p, 0u), n01 = ofdg_noise3(k0, k1, p + 1, 0u), n02 = ofdg_noise3(k0, k1, p + 2, 0u), n03 = ofdg_noise3(k0, k1, p + 3, 0u);
  const uint32_t n10 = ofdg_noise3(k0, k1, p, 1u), n11 = ofdg_noise3(k0, k1, p + 1, 1u), n12 = ofdg_noise3(k0, k1, p + 2, 1u), n13 = ofdg_noise3(k0, k1, p + 3, 1u);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 v0, v1;
    v0.x = ofdg_augment_apply(au, byte_to_float(a0, c), c, n00); v0.y = ofdg_augment_apply(au, byte_to_float(a1, c), c, n01);
    v0.z = ofdg_augment_apply(au, byte_to_float(a2, c), c, n02); v0.w = ofdg_augment_apply(au, byte_to_float(a3, c), c, n03);
    v1.x = ofdg_augment_apply(au, byte_to_float(b0, c), c, n10); v1.y = ofdg_augment_apply(au, byte_to_float(b1, c), c, n11);
    v1.z = ofdg_augment_apply(au, byte_to_float(b2, c), c, n12); v1.w = ofdg_augment_apply(au, byte_to_float(b3, c), c, n13);
    __stcs(reinterpret_cast<float4*>(o0 + c * P), v0);
    __stcs(reinterpret_cast<float4*>(o1 + c * P), v1);
  }
}

// ------------------------------------------------------------------------------------------------
// mode 9: non-rigid warp fields (consumer side, DG.cpp:237-252, 370-386, 403-406, 714-717)
// ------------------------------------------------------------------------------------------------
// CImg linear_atXY(fx, fy, z, c, out_value = 0) tap coordinates: x = (int)fx - (fx >= 0 ? 0 : 1) (not a
// true floor for negative integers -- kept). Returns false when every tap is out of range anyway
// (NaN / huge coordinates; the reference then produces NaN or 0, both of which store 0).
__device__ __forceinline__ bool dirichlet_setup(float fx, float fy, int& ix, int& iy, float& dx, float& dy) {
  if (!(fabsf(fx) < 1.0e9f) || !(fabsf(fy) < 1.0e9f)) return false;
  ix = (int)fx - (fx >= 0 ? 0 : 1);
  iy = (int)fy - (fy >= 0 ? 0 : 1);
  dx = fx - ix;
  dy = fy - iy;
  return true;
}
__device__ __forceinline__ float cimg_lerp2(float Icc, float Inc, float Icn, float Inn, float dx, float dy) {
  return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
}
__device__ __forceinline__ unsigned dirichlet_u8(const uint8_t* img, int w, int h, float fx, float fy) {
  int ix, iy;
  float dx, dy;
  if (!dirichlet_setup(fx, fy, ix, iy, dx, dy)) return 0u;
  auto tap = [&](int x, int y) -> float { return (x < 0 || y < 0 || x >= w || y >= h) ? 0.f : (float)img[(size_t)y * w + x]; };
  const float v = cimg_lerp2(tap(ix, iy), tap(ix + 1, iy), tap(ix, iy + 1), tap(ix + 1, iy + 1), dx, dy);
  return (unsigned)(unsigned char)v;
}
// Same interpolation over three packed channels; taps come from a callable (px, py) -> RGBX, 0 outside.
template <class Tap>
__device__ __forceinline__ uint32_t dirichlet_rgbx(Tap tap, float fx, float fy) {
  int ix, iy;
  float dx, dy;
  if (!dirichlet_setup(fx, fy, ix, iy, dx, dy)) return 0u;
  const uint32_t pcc = tap(ix, iy), pnc = tap(ix + 1, iy), pcn = tap(ix, iy + 1), pnn = tap(ix + 1, iy + 1);
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = cimg_lerp2(byte_to_float(pcc, c), byte_to_float(pnc, c), byte_to_float(pcn, c), byte_to_float(pnn, c), dx, dy);
    out |= ((uint32_t)(unsigned char)v) << (8 * c);
  }
  return out;
}
// applyWarpFieldToTexture over getTransformedTexture (DG.cpp:341-345, 670-681): the Dirichlet-bilinear sample at (fx, fy) of the
// sw x sh image that AGG's span filter would produce from `img` under the inverse matrix tinv -- each of its four taps is one
// span-bilinear pixel. The two taps of a row share the row's span interpolator (its set-up is two matrix products and four
// roundings in double), so it is built once per row, not once per tap.
__device__ __forceinline__ uint32_t warped_dirichlet_rgbx(const uchar4* img, int pitch, int sw, int sh, const double* tinv, float fx, float fy) {
  int ix, iy;
  float dx, dy;
  if (!dirichlet_setup(fx, fy, ix, iy, dx, dy)) return 0u;
  uint32_t px[4] = {0u, 0u, 0u, 0u};  // cc, nc, cn, nn; 0 outside
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int py = iy + r;
    if (py < 0 || py >= sh || ix + 1 < 0 || ix >= sw) continue;
    RowWarp rw;
    rw.init(tinv, (double)py, sw);
    if (ix >= 0) px[2 * r] = bilinear_rgbx(img, pitch, 0, 0, sw, sh, rw, ix);
    if (ix + 1 < sw) px[2 * r + 1] = bilinear_rgbx(img, pitch, 0, 0, sw, sh, rw, ix + 1);
  }
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = cimg_lerp2(byte_to_float(px[0], c), byte_to_float(px[1], c), byte_to_float(px[2], c), byte_to_float(px[3], c), dx, dy);
    out |= ((uint32_t)(unsigned char)v) << (8 * c);
  }
  return out;
}
// CImg _linear_atXY (Neumann) over a callable (x, y) -> float
template <class At>
__device__ __forceinline__ float neumann_f(At at, int w, int h, float fx, float fy) {
  const float nfx = fx <= 0 ? 0 : (fx >= w - 1 ? (float)(w - 1) : fx), nfy = fy <= 0 ? 0 : (fy >= h - 1 ? (float)(h - 1) : fy);
  const unsigned int x = (unsigned int)nfx, y = (unsigned int)nfy;
  const float dx = nfx - x, dy = nfy - y;
  const unsigned int nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
  return cimg_lerp2(at(x, y), at(nx, y), at(x, ny), at(nx, ny), dx, dy);
}
// Value at integer (X, Y) of a (W+1)x(H+1) field plane after CImg resize(2W, 2H, linear) and `*= 2.`
// (DG.cpp:1197-1200): x pass then y pass, each in double, each stored as float.
__device__ __forceinline__ float resized_field2(const float* f, int fw, int fh, int X, int Y, const RenderArgs& a) {
  const int px = a.fpos_x[X], py = a.fpos_y[Y];
  const double ax = a.falpha_x[X], ay = a.falpha_y[Y];
  const int px2 = px < fw - 1 ? px + 1 : px, py2 = py < fh - 1 ? py + 1 : py;
  const float r0 = (float)((1 - ax) * (double)f[(size_t)py * fw + px] + ax * (double)f[(size_t)py * fw + px2]);
  const float r1 = (float)((1 - ax) * (double)f[(size_t)py2 * fw + px] + ax * (double)f[(size_t)py2 * fw + px2]);
  const float v = (float)((1 - ay) * (double)r0 + ay * (double)r1);
  return (float)(v * 2.);
}

__device__ __forceinline__ bool box_hits_tile(const int32_t* b, int tx0, int ty0) {
  // cells right of the tile never matter; cells left of it feed the carry-in
  return b[1] <= ty0 + TH - 1 && b[3] >= ty0 && b[0] <= tx0 + TW - 1 && b[2] >= tx0;
}

// Mode 9: the pre-pass materialises a warped outline's frame-1 masks only where they can be non-zero -- the outline's
// frame-1 box (already widened by the field's reach, flatten.cpp) clipped to the frame, columns rounded outwards to whole
// four-pixel words. Everything outside reads as 0.
struct WarpRegion {
  int x0, y0, x1, y1;  // inclusive; x0 % 4 == 0, (x1 + 1) % 4 == 0 or x1 == W - 1
  __device__ __forceinline__ bool empty() const { return x1 < x0 || y1 < y0; }
};
__device__ __forceinline__ WarpRegion warp_region(const int32_t* b, int W, int H) {
  WarpRegion r;
  r.x0 = max(b[0], 0) & ~3; r.y0 = max(b[1], 0);
  r.x1 = min(min(b[2], W - 1) | 3, W - 1); r.y1 = min(b[3], H - 1);
  return r;
}
__device__ __forceinline__ void warped_mask_words(const RenderArgs& a, int slot, int x0, int y, uint32_t& aa, uint32_t& na) {
  const WarpRegion r = warp_region(a.shapes[a.deform_shape[slot]].bbox[1], a.W, a.H);
  if (x0 < r.x0 || x0 > r.x1 || y < r.y0 || y > r.y1) { aa = 0u; na = 0u; return; }
  const size_t P = (size_t)a.W * a.H;
  const uint8_t* mw = a.mask_warp + (size_t)slot * 2 * P + (size_t)y * a.W + x0;
  aa = *reinterpret_cast<const uint32_t*>(mw);
  na = *reinterpret_cast<const uint32_t*>(mw + P);
}

// ------------------------------------------------------------------------------------------------
// binning: which objects touch which tile (so that background-only tiles never enter the object path)
// ------------------------------------------------------------------------------------------------
__global__ void bin_kernel(RenderArgs a) {
  __shared__ int s_box[256][8];
  const int sample = blockIdx.x;
  const FlatSample& smp = a.samples[sample];
  const int n_obj = smp.obj_count;
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH, n_tiles = tiles_x * tiles_y;
  for (int i = threadIdx.x; i < min(n_obj, 256) * 8; i += blockDim.x) s_box[i >> 3][i & 7] = (&a.objects[smp.obj_begin + (i >> 3)].bbox[0][0])[i & 7];
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const int tx0 = (t % tiles_x) * TW, ty0 = (t / tiles_x) * TH;
    uint8_t* out = a.tile_hits + ((size_t)sample * n_tiles + t) * TILE_HIT_STRIDE;
    int cnt = 0;
    for (int o = 0; o < n_obj; ++o) {
      if (box_hits_tile(&s_box[o][0], tx0, ty0) || box_hits_tile(&s_box[o][4], tx0, ty0)) {
        if (cnt < TILE_HIT_STRIDE - 1) out[1 + cnt] = (uint8_t)o;
        ++cnt;
      }
    }
    out[0] = (uint8_t)(cnt <= TILE_HIT_STRIDE - 1 ? cnt : 255);
  }
}

// ------------------------------------------------------------------------------------------------
// render kernel: one CTA per (tile, sample); warp = tile row, lane = 4 consecutive pixels
// ------------------------------------------------------------------------------------------------
constexpr int NLAYER = 4;      // cover/area accumulator layers that are filled between two barriers
constexpr int MAX_HITS = 64;   // objects touching the tile handled per pass
constexpr int MAX_JOBS = 128;  // shapes (outlines) handled per pass
constexpr int MAX_PAIRS = 1024;  // (edge, tile row) work items listed per chunk; the rest is handled in place

struct HitObject {             // what the per-pixel stage needs of a FlatObject, staged in shared memory
  int obj;                     // index within the sample (z-order)
  int shape_begin, shape_count;
  int fg_pitch;                // foreground view of the object's texture (TexInfo): row pitch in pixels,
  unsigned long long fg_base;  //   first pixel in the pool
  int composite;
  int field;                   // mode 9 field id or -1
};
struct Job {                   // one outline of a hit object
  int vbegin[2], vcount[2];
  short hit;                   // index into the hit table
  signed char slot[2];         // accumulator layer per frame, -1: the outline misses the tile (coverage 0)
  unsigned char flags;         // 1 additive | 2 first outline of its object | 4 last outline | 8 a chunk ends after this job
  int deform;                  // mode 9: frame-1 masks come from this slot of the warped-mask scratch, else -1
};

// MovingObjectComposite::renderMasks (DG.cpp:606, 626) with the trivial cases folded: the strict-float
// formulas give ADD[u][255] = 255, SUB[u][255] = 0, SUB[0][v] = 0 and ADD[0][0] = 0 exactly; everything
// else goes through the float expression (u/255.f and v/255.f come from a table of those quotients).
__device__ __forceinline__ unsigned comp_add(unsigned u, unsigned v, const float* q255) {
  if (v == 255u) return 255u;
  if ((u | v) == 0u) return 0u;
  return (unsigned)(unsigned char)(255.f * (1.f - (1.f - q255[u]) * (1.f - q255[v])));
}
__device__ __forceinline__ unsigned comp_sub(unsigned u, unsigned v, const float* q255) {
  if (v == 255u || u == 0u) return 0u;
  return (unsigned)(unsigned char)(255.f * ((q255[u]) * (1.f - q255[v])));
}
// The same rules on four pixels packed one byte each (out of line: only outline pixels of composites get here).
__device__ __noinline__ uint32_t comp4_bytes(uint32_t u, uint32_t v, bool additive, const float* q255) {
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned ub = (u >> (8 * i)) & 255u, vb = (v >> (8 * i)) & 255u;
    out |= (additive ? comp_add(ub, vb, q255) : comp_sub(ub, vb, q255)) << (8 * i);
  }
  return out;
}
__device__ __forceinline__ uint32_t comp4(uint32_t u, uint32_t v, bool additive, const float* q255) {
  if (additive) {
    if (v == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    if ((u | v) == 0u) return 0u;
  } else {
    if (v == 0xFFFFFFFFu || u == 0u) return 0u;
  }
  return comp4_bytes(u, v, additive, q255);
}

// The same rules through the table composite_lut_kernel fills once per generator with exactly these float expressions
// (RenderArgs::comp_lut): four byte loads instead of four float evaluations behind per-byte branches. The raster kernel
// spent a fifth of its instructions in comp4_bytes, at nine active lanes per instruction.
__device__ __forceinline__ uint32_t comp4_lut(uint32_t u, uint32_t v, bool additive, const uint8_t* lut) {
  if (additive) {
    if (v == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    if ((u | v) == 0u) return 0u;
  } else {
    if (v == 0xFFFFFFFFu || u == 0u) return 0u;
  }
  const uint8_t* t = lut + (additive ? 0 : 65536);
  const uint32_t lo = __byte_perm(v, u, 0x5140u), hi = __byte_perm(v, u, 0x7362u);  // {v0, u0, v1, u1}, {v2, u2, v3, u3}: index = u * 256 + v
  const uint32_t r0 = __ldg(t + (lo & 0xFFFFu)), r1 = __ldg(t + (lo >> 16)), r2 = __ldg(t + (hi & 0xFFFFu)), r3 = __ldg(t + (hi >> 16));
  return __byte_perm(__byte_perm(r0, r1, 0x0040u), __byte_perm(r2, r3, 0x0040u), 0x5410u);
}

// kDeform = false compiles the mode-9 (warp field) branches out; kExtra = false the extra tops (backward flow, ids).
template <bool kDeform, bool kExtra>
__global__ void __launch_bounds__(RENDER_THREADS, OFDG_RENDER_MIN_BLOCKS) render_kernel(RenderArgs a) {
  __shared__ int s_cover[NLAYER][TH][TW];
  __shared__ int s_area[NLAYER][TH][TW];
  __shared__ int s_carry[NLAYER][TH];
  __shared__ float s_q255[256];
  __shared__ HitObject s_hit[MAX_HITS];
  __shared__ Job s_job[MAX_JOBS];
  __shared__ int s_jobbase[MAX_HITS + 1];
  __shared__ int s_njob, s_hits_done, s_next_obj;
  __shared__ int s_seg_begin[NLAYER], s_seg_count[NLAYER];
  __shared__ unsigned s_pairs[MAX_PAIRS];  // (layer, tile row, edge) work items of the current chunk
  __shared__ int s_npairs;

  const int W = a.W, H = a.H;
  const int tiles_x = (W + TW - 1) / TW;
  const int tx0 = (blockIdx.x % tiles_x) * TW, ty0 = (blockIdx.x / tiles_x) * TH;
  const int sample = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int y = ty0 + warp, x0 = tx0 + lane * 4;
  const bool live = (y < H) && (x0 < W);
  const FlatSample& smp = a.samples[sample];
  const size_t P = (size_t)W * H;
  const int n_obj = smp.obj_count, obj_begin = smp.obj_begin;

  const uint8_t* bins = a.tile_hits + ((size_t)sample * gridDim.x + blockIdx.x) * TILE_HIT_STRIDE;
  const int binned = bins[0];  // objects touching this tile (255: more than a bin entry lists)

  if (binned)  // the composite rules' quotient table: only tiles with objects can need it
    for (int i = tid; i < 256; i += RENDER_THREADS) s_q255[i] = (float)i / 255.f;

  // ---- pass set-up, entirely inside warp 0 (the other warps fetch the background meanwhile):
  //      hit table -> outline jobs -> accumulator layers and chunk boundaries
  auto pass_setup = [&](int obj0) {
    // (1) objects whose boxes touch the tile, in z-order: from the bin list, or by a scan when the tile
    //     holds more objects than a bin entry can list
    int nh = 0, next = n_obj;
    if (binned <= TILE_HIT_STRIDE - 1) {
      const int idx = lane < binned ? bins[1 + lane] : 0x7FFFFFFF;
      // a later pass (more outlines than one pass holds) resumes behind the objects already done
      const int skip = __popc(__ballot_sync(0xffffffffu, idx < obj0));
      if (lane < binned && idx >= obj0) {
        const FlatObject* ob = a.objects + obj_begin + idx;
        HitObject h;
        const TexInfo& ti = a.tex_info[ob->tex];
        h.obj = idx; h.shape_begin = ob->shape_begin; h.shape_count = ob->shape_count;
        h.fg_pitch = ti.fg_pitch; h.fg_base = ti.fg_base;
        h.composite = ob->composite; h.field = ob->field;
        s_hit[lane - skip] = h;
      }
      nh = binned - skip;
    } else {
      for (int o = obj0; o < n_obj; o += 32) {
        const int idx = o + lane;
        bool hit = false;
        const FlatObject* ob = a.objects + obj_begin + idx;
        if (idx < n_obj) hit = box_hits_tile(ob->bbox[0], tx0, ty0) || box_hits_tile(ob->bbox[1], tx0, ty0);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        const int pos = nh + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < MAX_HITS) {
          HitObject h;
          const TexInfo& ti = a.tex_info[ob->tex];
          h.obj = idx; h.shape_begin = ob->shape_begin; h.shape_count = ob->shape_count;
          h.fg_pitch = ti.fg_pitch; h.fg_base = ti.fg_base;
          h.composite = ob->composite; h.field = ob->field;
          s_hit[pos] = h;
        }
        const int cnt = __popc(bal);
        if (nh + cnt > MAX_HITS) {  // table full: the next pass rescans from the first object left out
          next = o + (int)__fns(bal, 0, MAX_HITS - nh + 1);
          nh = MAX_HITS;
          break;
        }
        nh += cnt;
      }
    }
    __syncwarp();
    // (2) job ranges of the hit objects (at most MAX_JOBS outlines per pass)
    if (lane == 0) {
      int nj = 0, h = 0;
      for (; h < nh; ++h) {
        if (nj + s_hit[h].shape_count > MAX_JOBS && h > 0) break;
        s_jobbase[h] = nj;
        nj += min(s_hit[h].shape_count, MAX_JOBS);
      }
      s_jobbase[h] = nj;
      s_hits_done = h;
      s_njob = nj;
      s_next_obj = (h == nh) ? next : s_hit[h].obj;  // first object not handled by this pass
    }
    __syncwarp();
    const int njob = s_njob, nhd = s_hits_done;
    for (int jj = lane; jj < njob; jj += 32) {
      int h = 0;
      while (h + 1 < nhd && s_jobbase[h + 1] <= jj) ++h;
      const int si = jj - s_jobbase[h];
      const HitObject ho = s_hit[h];
      const FlatShape& sh = a.shapes[ho.shape_begin + si];
      Job j;
      j.hit = (short)h;
      const bool h0 = box_hits_tile(sh.bbox[0], tx0, ty0), h1 = box_hits_tile(sh.bbox[1], tx0, ty0);
      j.vbegin[0] = sh.vbegin[0]; j.vbegin[1] = sh.vbegin[1];
      j.deform = (kDeform && h1) ? sh.deform : -1;  // a warped outline's frame-1 masks were materialised by the pre-pass
      const bool r1 = h1 && j.deform < 0;
      j.vcount[0] = h0 ? sh.vcount[0] : 0; j.vcount[1] = r1 ? sh.vcount[1] : 0;
      j.slot[0] = h0 ? 0 : -1; j.slot[1] = r1 ? 0 : -1;
      j.flags = (unsigned char)((sh.additive ? 1 : 0) | (si == 0 ? 2 : 0) | (si == ho.shape_count - 1 ? 4 : 0));
      s_job[jj] = j;
    }
    __syncwarp();
    // (3) accumulator layers and chunk boundaries
    if (lane == 0 && njob > 0) {
      int used = 0;
      for (int j = 0; j < njob; ++j) {
        const int need = (s_job[j].slot[0] >= 0) + (s_job[j].slot[1] >= 0);
        if (used + need > NLAYER) { s_job[j - 1].flags |= 8; used = 0; }
        if (s_job[j].slot[0] >= 0) s_job[j].slot[0] = (signed char)used++;
        if (s_job[j].slot[1] >= 0) s_job[j].slot[1] = (signed char)used++;
      }
      s_job[njob - 1].flags |= 8;
    }
  };

  int obj0 = binned ? 0 : n_obj;  // background-only tiles skip the object path entirely
  if (obj0 < n_obj && warp == 0) pass_setup(obj0);

  uint32_t col0[4], col1[4];
  uint32_t id0 = 0, id1 = 0;  // four pixels, one byte each: 0 = background, k+1 = k-th foreground object (k < 255)

  // ---- background: masks are all 255 (DG.cpp:684-690); frame 0 = centre window of the prepared
  //      texture, frame 1 = that texture warped by I^-1*M*I on the 2W x 2H canvas (DG.cpp:665-682)
  {
    const uchar4* bg = a.bg + (size_t)sample * (4 * P);
    const int W2 = 2 * W, H2 = 2 * H;
    if (live) {
      const uchar4* row = bg + (size_t)(y + H / 2) * W2 + (x0 + W / 2);
      RowWarp rw;
      rw.init(smp.bg_tex_inv, (double)(y + H / 2), W2);
#pragma unroll
      for (int i = 0; i < 4; ++i) col0[i] = ld_px(row + i) & 0xFFFFFFu;
#pragma unroll
      for (int i = 0; i < 4; ++i) col1[i] = bilinear_rgbx_call(bg, W2, W2, H2, W2, rw.dx.v1, rw.dx.lft, rw.dx.rem, rw.dy.v1, rw.dy.lft, rw.dy.rem, x0 + i + W / 2);
      if (kDeform && smp.bg_field >= 0) {
        // background with a warp field: the warped 2W x 2H texture is resampled through the resized,
        // doubled inverse field before the centre crop (DG.cpp:670-681, 1194-1201)
        const int fw = W + 1, fh = H + 1;
        const float* ifl = a.fields + ((size_t)smp.bg_field * 2 + 1) * 2 * fw * fh;
        for (int i = 0; i < 4; ++i) {
          const int X = x0 + i + W / 2, Y = y + H / 2;
          const float sx = X + resized_field2(ifl, fw, fh, X, Y, a), sy = Y + resized_field2(ifl + (size_t)fw * fh, fw, fh, X, Y, a);
          col1[i] = warped_dirichlet_rgbx(bg, W2, W2, H2, smp.bg_tex_inv, sx, sy);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) col0[i] = col1[i] = 0;
    }
  }

  uint32_t uaa[2] = {0, 0}, una[2] = {0, 0};  // masks of the object being assembled: [frame], 4 pixels x 1 byte

  // ---- foreground objects in z-order, in passes of at most MAX_HITS objects / MAX_JOBS outlines
  while (obj0 < n_obj) {
    __syncthreads();  // warp 0's set-up (and s_q255) are visible
    const int njob = s_njob;

    // (3) chunks: zero -> accumulate edges -> per-pixel masks, combine, blit
    int j0 = 0;
    while (j0 < njob) {
      int j1 = j0;
      while (!(s_job[j1].flags & 8)) ++j1;
      ++j1;  // jobs [j0, j1)
      if (tid < NLAYER) {  // thread l publishes the edge list that feeds accumulator layer l
        int b = 0, c = 0;
        for (int j = j0; j < j1; ++j)
          for (int f = 0; f < 2; ++f)
            if (s_job[j].slot[f] == tid) { b = s_job[j].vbegin[f]; c = s_job[j].vcount[f]; }
        s_seg_begin[tid] = b;
        s_seg_count[tid] = c;
      }
      for (int i = tid; i < NLAYER * TH * TW / 4; i += RENDER_THREADS) {
        reinterpret_cast<int4*>(&s_cover[0][0][0])[i] = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4*>(&s_area[0][0][0])[i] = make_int4(0, 0, 0, 0);
      }
      if (tid < NLAYER * TH) (&s_carry[0][0])[tid] = 0;
      if (tid == 0) s_npairs = 0;
      __syncthreads();
      {
        // (a) threads over edges: which tile rows does the edge cross? One work item per (edge, row).
        const int c0 = s_seg_count[0], c1 = c0 + s_seg_count[1], c2 = c1 + s_seg_count[2], c3 = c2 + s_seg_count[3];
        for (int e = tid; e < c3; e += RENDER_THREADS) {
          const int l = e < c0 ? 0 : (e < c1 ? 1 : (e < c2 ? 2 : 3));
          const int ei = e - (l == 0 ? 0 : (l == 1 ? c0 : (l == 2 ? c1 : c2)));
          const int n = s_seg_count[l];
          const FlatVertex* v = a.verts + s_seg_begin[l];
          const FlatVertex p = v[ei], q = v[ei + 1 == n ? 0 : ei + 1];
          int rlo, rhi;
          bool left;
          if (!tile_edge_rows(tx0, ty0, p.x, p.y, q.x, q.y, rlo, rhi, left)) continue;
          const int nrows = rhi - rlo + 1;
          const int base = left ? MAX_PAIRS : atomicAdd(&s_npairs, nrows);
          for (int k = 0; k < nrows; ++k) {
            if (base + k < MAX_PAIRS) s_pairs[base + k] = ((unsigned)l << 28) | ((unsigned)(rlo + k - ty0) << 24) | (unsigned)ei;
            else tile_edge_row<true>(&s_cover[l][0][0], &s_area[l][0][0], &s_carry[l][0], tx0, ty0, rlo + k, p.x, p.y, q.x, q.y);  // cheap (left of the tile) or list full
          }
        }
      }
      __syncthreads();
      {
        // (b) threads over (edge, row) items: closed-form row segment -> cells
        const int np = min(s_npairs, MAX_PAIRS);
        for (int i = tid; i < np; i += RENDER_THREADS) {
          const unsigned w = s_pairs[i];
          const int l = (int)(w >> 28), r = ty0 + (int)((w >> 24) & 15u), ei = (int)(w & 0xFFFFFFu);
          const int n = s_seg_count[l];
          const FlatVertex* v = a.verts + s_seg_begin[l];
          const FlatVertex p = v[ei], q = v[ei + 1 == n ? 0 : ei + 1];
          tile_edge_row<true>(&s_cover[l][0][0], &s_area[l][0][0], &s_carry[l][0], tx0, ty0, r, p.x, p.y, q.x, q.y);
        }
      }
      __syncthreads();
      for (int j = j0; j < j1; ++j) {
        const Job jb = s_job[j];
        uint32_t vaa[2] = {0, 0}, vna[2] = {0, 0};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const int l = jb.slot[f];
          if (l < 0) {
            if (kDeform && f == 1 && jb.deform >= 0 && live) warped_mask_words(a, jb.deform, x0, y, vaa[1], vna[1]);
            continue;
          }
          const int4 c4 = *reinterpret_cast<const int4*>(&s_cover[l][warp][lane * 4]);
          const int4 a4 = *reinterpret_cast<const int4*>(&s_area[l][warp][lane * 4]);
          const int carry = s_carry[l][warp];
          const bool cells = (c4.x | c4.y | c4.z | c4.w | a4.x | a4.y | a4.z | a4.w) != 0;
          if (!__any_sync(0xffffffffu, cells)) {
            // no outline crosses this row inside the tile: coverage is constant along it
            if (carry == 0) continue;
            const int cv = coverage_alpha(carry, 0);
            vaa[f] = graylut((unsigned)cv) * 0x01010101u;
            vna[f] = cv >= 128 ? 0xFFFFFFFFu : 0u;
            continue;
          }
          int c[4] = {c4.x, c4.y, c4.z, c4.w}, ar[4] = {a4.x, a4.y, a4.z, a4.w};
          c[1] += c[0]; c[2] += c[1]; c[3] += c[2];
          int tot = c[3];  // warp-level inclusive prefix sum over the lanes' cover totals
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, tot, d);
            if (lane >= d) tot += o;
          }
          const int base = tot - c[3] + carry;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int cv = coverage_alpha(base + c[i], ar[i]);
            vaa[f] |= graylut((unsigned)cv) << (8 * i);        // gamma_none
            vna[f] |= (cv >= 128 ? 255u : 0u) << (8 * i);      // gamma_threshold(0.5), then graylut(255) = 255
          }
        }
        const HitObject ho = s_hit[jb.hit];
        if (ho.composite) {
          if (jb.flags & 2) { uaa[0] = uaa[1] = una[0] = una[1] = 0; }
          const bool add = jb.flags & 1;
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            uaa[f] = comp4(uaa[f], vaa[f], add, s_q255);
            // non-AA masks stay in {0, 255} (the rules are closed on it) unless a warp field resampled them
            if (kDeform) una[f] = comp4(una[f], vna[f], add, s_q255);
            else una[f] = add ? (una[f] | vna[f]) : (una[f] & ~vna[f]);
          }
        } else {
          uaa[0] = vaa[0]; uaa[1] = vaa[1]; una[0] = vna[0]; una[1] = vna[1];
        }
        if (!(jb.flags & 4) || !live) continue;

        // the object's masks are complete: ids from the non-AA masks, colour through the AA (or non-AA) masks
        const int k = ho.obj;
        if (a.dbg_masks && k < a.dbg_max_objs) {
          uint8_t* mb = a.dbg_masks + ((size_t)sample * a.dbg_max_objs + k) * 4 * P + (size_t)y * W + x0;
          *reinterpret_cast<uint32_t*>(mb + 0 * P) = uaa[0]; *reinterpret_cast<uint32_t*>(mb + 1 * P) = uaa[1];
          *reinterpret_cast<uint32_t*>(mb + 2 * P) = una[0]; *reinterpret_cast<uint32_t*>(mb + 3 * P) = una[1];
        }
        const uint32_t kk = (uint32_t)(k + 1) * 0x01010101u;
        const uint32_t e0 = __vcmpeq4(una[0], 0xFFFFFFFFu), e1 = __vcmpeq4(una[1], 0xFFFFFFFFu);
        id0 = (id0 & ~e0) | (kk & e0);
        id1 = (id1 & ~e1) | (kk & e1);
        const uint32_t m0w = a.use_aa ? uaa[0] : una[0], m1w = a.use_aa ? uaa[1] : una[1];
        if ((m0w | m1w) == 0u) continue;
        const uchar4* tex = a.pool + ho.fg_base;  // the W x H foreground view (centre crop, DG.cpp:99-102 with defaults, or the resized copy)
        if (m0w) {
          const uchar4* trow = tex + (size_t)y * ho.fg_pitch + x0;  // identity warp == copy
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned m0 = (m0w >> (8 * i)) & 255u;
            if (m0) col0[i] = blend_rgbx(col0[i], ld_px(trow + i) & 0xFFFFFFu, m0);
          }
        }
        if (m1w && (!kDeform || ho.field < 0)) {
          RowWarp rw;
          rw.init(a.objects[obj_begin + k].tex_inv, (double)y, W);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned m1 = (m1w >> (8 * i)) & 255u;
            if (m1) col1[i] = blend_rgbx(col1[i], bilinear_rgbx_call(tex, ho.fg_pitch, W, H, W, rw.dx.v1, rw.dx.lft, rw.dx.rem, rw.dy.v1, rw.dy.lft, rw.dy.rem, x0 + i), m1);
          }
        } else if (kDeform && m1w) {
          // applyWarpFieldToTexture(getTransformedTexture(tex0, M), iflow) evaluated where the mask is set:
          // each of the 4 float-bilinear taps is itself one AGG span-bilinear pixel (DG.cpp:341-345)
          const double* tinv = a.objects[obj_begin + k].tex_inv;
          const int fw = W + 1, fh = H + 1;
          const float* ifl = a.fields + ((size_t)ho.field * 2 + 1) * 2 * fw * fh;
          auto tap = [&](int px, int py) -> uint32_t {
            if (px < 0 || py < 0 || px >= W || py >= H) return 0u;
            RowWarp rw;
            rw.init(tinv, (double)py, W);
            return bilinear_rgbx(tex, ho.fg_pitch, 0, 0, W, H, rw, px);
          };
          for (int i = 0; i < 4; ++i) {
            const unsigned m1 = (m1w >> (8 * i)) & 255u;
            if (!m1) continue;
            const int x = x0 + i;
            const float sx = x + ifl[(size_t)y * fw + x], sy = y + ifl[(size_t)fw * fh + (size_t)y * fw + x];
            col1[i] = blend_rgbx(col1[i], dirichlet_rgbx(tap, sx, sy), m1);
          }
        }
      }
      j0 = j1;
      __syncthreads();  // the accumulators are rewritten by the next chunk
    }
    obj0 = s_next_obj;
    if (obj0 >= n_obj) break;
    __syncthreads();  // everybody has read the pass state before warp 0 rewrites it
    if (warp == 0) pass_setup(obj0);
  }

  if (!live) return;

  // ---- flow of the top-most object, f64 -> f32 (DG.cpp:388-401, 692-712). Forward: frame 0's ids through the motions;
  //      backward (extra top, computeFlowImage(inverse = true)): frame 1's ids through the inverse motions.
  auto point_flow = [&](unsigned oid, bool inverse, int i, float& fx, float& fy) {
    const float xf = (float)(x0 + i), yf = (float)y;
    // background: the point goes through I^-1 = T(-W,-H), M, I = T(W,H) (DG.cpp:697-712); objects: through M alone.
    // One code path: the translations are exact no-ops (+-0.0) for objects.
    const FlatObject* fo = oid ? a.objects + obj_begin + oid - 1 : nullptr;
    const double* m = oid ? (inverse ? fo->tex_inv : fo->motion) : (inverse ? smp.bg_motion_inv : smp.bg_motion);
    const double pre_x = oid ? 0.0 : (double)W, pre_y = oid ? 0.0 : (double)H;
    const float save_x = oid ? xf : xf + (float)(W / 2), save_y = oid ? yf : yf + (float)(H / 2);
    double ix = (double)save_x - pre_x, iy = (double)save_y - pre_y;
    const double tmp = ix;
    ix = tmp * m[0] + iy * m[2] + m[4];
    iy = tmp * m[1] + iy * m[3] + m[5];
    ix = ix + pre_x; iy = iy + pre_y;
    fx = (float)(ix - save_x);
    fy = (float)(iy - save_y);
    if (kDeform) {  // the forward field is added in both directions (DG.cpp:403-406, 714-717)
      const int fw = W + 1, fh = H + 1;
      if (oid == 0) {
        if (smp.bg_field >= 0 && ix >= 0 && ix < 2 * W && iy >= 0 && iy < 2 * H) {  // DG.cpp:714-717
          const float* fl = a.fields + ((size_t)smp.bg_field * 2 + 0) * 2 * fw * fh;
          auto at0 = [&](unsigned X, unsigned Y) { return resized_field2(fl, fw, fh, (int)X, (int)Y, a); };
          auto at1 = [&](unsigned X, unsigned Y) { return resized_field2(fl + (size_t)fw * fh, fw, fh, (int)X, (int)Y, a); };
          fx += neumann_f(at0, 2 * W, 2 * H, (float)ix, (float)iy);
          fy += neumann_f(at1, 2 * W, 2 * H, (float)ix, (float)iy);
        }
      } else if (fo->field >= 0 && ix >= 0 && ix < W && iy >= 0 && iy < H) {  // DG.cpp:403-406
        const float* fl = a.fields + ((size_t)fo->field * 2 + 0) * 2 * fw * fh;
        auto at0 = [&](unsigned X, unsigned Y) { return fl[(size_t)Y * fw + X]; };
        auto at1 = [&](unsigned X, unsigned Y) { return fl[(size_t)fw * fh + (size_t)Y * fw + X]; };
        fx += neumann_f(at0, fw, fh, (float)ix, (float)iy);
        fy += neumann_f(at1, fw, fh, (float)ix, (float)iy);
      }
    }
  };
  float fxv[4], fyv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) point_flow((id0 >> (8 * i)) & 255u, false, i, fxv[i], fyv[i]);

  // ---- write the three blobs (NCHW float): 8 planes x one 128-bit store per lane
  const size_t pix = (size_t)y * W + x0;
  float* of = a.flow + (size_t)sample * 2 * P + pix;
  if (a.img0) {
    float* o0 = a.img0 + (size_t)sample * 3 * P + pix;
    float* o1 = a.img1 + (size_t)sample * 3 * P + pix;
    if (smp.aug.enabled != 0) {  // this repository's own colour/noise augmentation (ofdg/augment.h); never set by the reference path
      store_augmented(&smp.aug, col0[0], col0[1], col0[2], col0[3], col1[0], col1[1], col1[2], col1[3], o0, o1, P, (uint32_t)pix);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 v0 = make_float4(byte_to_float(col0[0], c), byte_to_float(col0[1], c), byte_to_float(col0[2], c), byte_to_float(col0[3], c));
        const float4 v1 = make_float4(byte_to_float(col1[0], c), byte_to_float(col1[1], c), byte_to_float(col1[2], c), byte_to_float(col1[3], c));
        __stcs(reinterpret_cast<float4*>(o0 + c * P), v0);
        __stcs(reinterpret_cast<float4*>(o1 + c * P), v1);
      }
    }
  }
  __stcs(reinterpret_cast<float4*>(of), make_float4(fxv[0], fxv[1], fxv[2], fxv[3]));
  __stcs(reinterpret_cast<float4*>(of + P), make_float4(fyv[0], fyv[1], fyv[2], fyv[3]));

  if (kExtra) {
    if (a.flow_bw) {
      float bx[4], by[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) point_flow((id1 >> (8 * i)) & 255u, true, i, bx[i], by[i]);
      float* ob = a.flow_bw + (size_t)sample * 2 * P + pix;
      __stcs(reinterpret_cast<float4*>(ob), make_float4(bx[0], bx[1], bx[2], bx[3]));
      __stcs(reinterpret_cast<float4*>(ob + P), make_float4(by[0], by[1], by[2], by[3]));
    }
    if (a.top_id0 || a.top_id1) {
      float v0[4], v1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned b0 = (id0 >> (8 * i)) & 255u, b1 = (id1 >> (8 * i)) & 255u;
        v0[i] = (float)(b0 ? a.objects[obj_begin + b0 - 1].obj_id : 1);  // the background's ID is 1 (data_generation_layer.cpp:199)
        v1[i] = (float)(b1 ? a.objects[obj_begin + b1 - 1].obj_id : 1);
      }
      if (a.top_id0) __stcs(reinterpret_cast<float4*>(a.top_id0 + (size_t)sample * P + pix), make_float4(v0[0], v0[1], v0[2], v0[3]));
      if (a.top_id1) __stcs(reinterpret_cast<float4*>(a.top_id1 + (size_t)sample * P + pix), make_float4(v1[0], v1[1], v1[2], v1[3]));
    }
    if (a.ids8) {
      *reinterpret_cast<uint32_t*>(a.ids8 + (size_t)sample * 2 * P + pix) = id0;
      *reinterpret_cast<uint32_t*>(a.ids8 + (size_t)sample * 2 * P + P + pix) = id1;
    }
  }

  if (a.dbg_id0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned b0 = (id0 >> (8 * i)) & 255u, b1 = (id1 >> (8 * i)) & 255u;
      const unsigned o0id = b0 ? (unsigned)a.objects[obj_begin + b0 - 1].obj_id : 1u;
      const unsigned o1id = b1 ? (unsigned)a.objects[obj_begin + b1 - 1].obj_id : 1u;
      a.dbg_id0[(size_t)sample * P + pix + i] = o0id;
      if (a.dbg_id1) a.dbg_id1[(size_t)sample * P + pix + i] = o1id;
    }
  }
  if (a.frames8) {  // byte planes: channel c of the lane's four pixels is one 32-bit store
    uint8_t* fb = a.frames8 + (size_t)sample * 6 * P + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned sel = 0x40u + (unsigned)c * 0x11u;  // byte c of the first operand, byte c of the second
      const uint32_t w0 = __byte_perm(__byte_perm(col0[0], col0[1], sel), __byte_perm(col0[2], col0[3], sel), 0x5410u);
      const uint32_t w1 = __byte_perm(__byte_perm(col1[0], col1[1], sel), __byte_perm(col1[2], col1[3], sel), 0x5410u);
      __stcs(reinterpret_cast<uint32_t*>(fb + c * P), w0);
      __stcs(reinterpret_cast<uint32_t*>(fb + (3 + c) * P), w1);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// split render path: bin -> raster (masks per (object, tile) pair) -> shade (background, blits, flow, blobs)
// ------------------------------------------------------------------------------------------------
// The fused kernel above keeps the colours of a tile in registers while it rasterises: 80 registers, 44 KB of shared
// memory and block-wide barriers for the whole tile. Here the two halves are separate kernels. The raster kernel works on
// one (object, tile) pair per block turn -- no z-order between pairs, so every pair of the batch is independent work -- and
// leaves the object's four masks (AA / non-AA, both frames; composites already combined) in HBM: 4 KB per pair, a few
// hundred pairs per sample. The shade kernel is barrier-free: each warp owns a tile row and blends the tile's pairs in
// z-order straight from those masks.
// The raster kernel's work unit is a horizontal slice of a pair's tile: RTH rows, one warp per row. Slices of a pair are
// independent (every (edge, row) contribution is closed-form), and smaller blocks mean more of them per SM: the phases of one
// unit are separated by block-wide barriers whose wait is set by the slowest thread, so many small blocks hide it better.
#ifndef OFDG_RASTER_ROWS
#define OFDG_RASTER_ROWS 8  // measured: 8 / 4 / 2 rows (5 / 10 / 20 blocks per SM) -> 0.160 / 0.167 / 0.191 ms
#endif
constexpr int RTH = OFDG_RASTER_ROWS;
constexpr int RSUB = TH / RTH;              // slices per pair
// Warps per raster block and accumulator layers per chunk. A unit's phases are dependent chains separated by block barriers
// (shape records -> vertices -> row crossings -> shared atomics -> sweep), and only the sweep has work for every thread. Four
// warps (each sweeping two rows) and one outline per chunk (two layers, 19 KB of shared memory) put eight units on an SM at 64
// registers instead of five at 48. Measured (same box, kernel in line / whole pipelined step): 8 warps x 4 layers x 5 blocks
// 0.153 / 0.4255 ms; 8 x 2 x 5: 0.171 / 0.435; 4 x 2 x 10 (48 registers, spills): 0.173 / 0.432; 4 x 2 x 8: 0.151 / 0.407;
// 4 x 2 x 6: 0.166 / 0.424; 4 x 4 x 5: 0.166 / 0.430; 2 x 2 x 11: 0.242 / 0.487. With the composite table (below): 4 x 2 x 7 / 8 / 9 /
// 10 blocks (72 / 64 / 56 / 48 registers) 0.138 / 0.130 / 0.138 / 0.138 ms. (Starting the packed item loops at a warp that
// changes from unit to unit, so that they do not all issue from the same scheduler: no gain, 0.153 vs 0.151.)
#ifndef OFDG_RASTER_WARPS
#define OFDG_RASTER_WARPS 4
#endif
#ifndef OFDG_RASTER_LAYERS
#define OFDG_RASTER_LAYERS 2
#endif

constexpr int RWARPS = OFDG_RASTER_WARPS;
constexpr int RPW = RTH / RWARPS;           // tile rows swept per warp
constexpr int RLAYER = OFDG_RASTER_LAYERS;  // accumulator layers of the raster kernel: RLAYER / 2 outlines per chunk
constexpr int RASTER_THREADS = 32 * RWARPS;
constexpr int RASTER_ITEMS = MAX_PAIRS * RTH * RLAYER / (TH * NLAYER);  // (edge, row) work items listed per chunk
static_assert(TH % RTH == 0, "a raster slice must divide the tile");
static_assert(RTH % RWARPS == 0 && RLAYER % 2 == 0 && RLAYER <= NLAYER && RLAYER * RTH <= RASTER_THREADS, "raster block shape");
#ifndef OFDG_RASTER_MIN_BLOCKS
#define OFDG_RASTER_MIN_BLOCKS (OFDG_RASTER_WARPS == 4 ? 8 : 1280 / (32 * OFDG_RASTER_WARPS))  // 8 warps, measured: 3 / 4 / 5 / 6 blocks per SM -> 0.225 / 0.195 / 0.179 / 0.188 ms
#endif
__device__ __forceinline__ bool box_hits_rows(const int32_t* b, int tx0, int ty0, int rows) {
  return b[1] <= ty0 + rows - 1 && b[3] >= ty0 && b[0] <= tx0 + TW - 1 && b[2] >= tx0;
}
#ifndef OFDG_SHADE_DEFORM_MIN_BLOCKS
#define OFDG_SHADE_DEFORM_MIN_BLOCKS 4  // the mode-9 instances (warp-field branches compiled in): 5 / 4 / 3 blocks per SM (48 / 64 / 80 registers) -> config 3 shade 0.473 / 0.450 / 0.452 ms
#endif
#ifndef OFDG_SHADE_MIN_BLOCKS
#define OFDG_SHADE_MIN_BLOCKS 5  // measured (round 2 kernel): 4 / 5 / 6 blocks per SM -> 0.152 / 0.142 / 0.153 ms
#endif
// ---- bulk asynchronous copies global -> shared (cp.async.bulk, completion counted on an mbarrier), sm_90+
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {  // 16-byte aligned, size a multiple of 16
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#ifndef OFDG_SHADE_STAGE
#define OFDG_SHADE_STAGE 0  // > 0: that many (object, tile) pairs' masks and records are brought into shared memory by one bulk copy (cp.async.bulk +
                            // mbarrier) while the background is filtered. Measured with 6: shade 0.142 -> 0.160 ms (5 blocks/SM), 0.153 -> 0.176 ms (6):
                            // the 26 KB per block come out of the L1 the bilinear taps live in. 0: per-pair global loads + L1 prefetch of the next pair.
#endif

constexpr int PAIR_ROW_STRIDE = 2 + 2 * TH;  // RenderArgs::pair_rows per pair: a two-word header {foreground view base (2 x 32 bits), pitch, object} {field}, then two words per tile row
struct PairOutline {   // one outline of the pair's object, staged in shared memory
  int vbegin[2], vcount[2];
  signed char layer[2];  // accumulator layer per frame, -1: the outline misses the tile
  unsigned char additive;
  int deform;
};

constexpr int BIN_THREADS = 64;  // one tile per thread; a sample's tiles are split over gridDim.y blocks
__global__ void __launch_bounds__(BIN_THREADS) bin_pairs_kernel(RenderArgs a) {
  __shared__ int s_box[256][8];
  __shared__ int2 s_shapes[256];  // per object: first outline, outline count | composite << 16 (saves the raster kernel two dependent loads per pair)
  __shared__ int s_scan[BIN_THREADS];
  __shared__ int s_base;
  // grid = (sample, part): the parts of a sample split its tiles (and its background rows) among them; each part claims its
  // own slice of the pair list, so the pairs of a tile stay consecutive, which is all the raster and shade kernels rely on
  const int sample = blockIdx.x, tid = threadIdx.x, part = blockIdx.y, n_parts = gridDim.y;
  const FlatSample& smp = a.samples[sample];
  const int n_obj = min(smp.obj_count, 255);
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH, n_tiles = tiles_x * tiles_y;
  for (int i = tid; i < n_obj * 8; i += blockDim.x) s_box[i >> 3][i & 7] = (&a.objects[smp.obj_begin + (i >> 3)].bbox[0][0])[i & 7];
  for (int o = tid; o < n_obj; o += blockDim.x) {
    const FlatObject& ob = a.objects[smp.obj_begin + o];
    s_shapes[o] = make_int2(ob.shape_begin, ob.shape_count | (ob.composite ? 1 << 16 : 0));
  }
  __syncthreads();
  if (a.bg_rows) {  // the background's span-interpolator rows (frame 1 = the prepared texture under I^-1 * M * I, DG.cpp:665-682)
    for (int y = tid + part * blockDim.x; y < a.H; y += blockDim.x * n_parts) {
      RowWarp rw;
      rw.init(smp.bg_tex_inv, (double)(y + a.H / 2), 2 * a.W);
      int4* r = a.bg_rows + ((size_t)sample * a.H + y) * 2;
      r[0] = make_int4(rw.dx.v1, rw.dx.lft, rw.dx.rem, rw.dy.v1);
      r[1] = make_int4(rw.dy.lft, rw.dy.rem, 0, 0);
    }
  }
  int mine = 0;  // pairs of this thread's tiles
  for (int t = tid + part * blockDim.x; t < n_tiles; t += blockDim.x * n_parts) {
    const int tx0 = (t % tiles_x) * TW, ty0 = (t / tiles_x) * TH;
    uint8_t* out = a.tile_hits + ((size_t)sample * n_tiles + t) * TILE_HIT_STRIDE;  // the fused kernel's bin entry (its overflow fallback)
    int cnt = 0;
    for (int o = 0; o < n_obj; ++o)
      if (box_hits_tile(&s_box[o][0], tx0, ty0) || box_hits_tile(&s_box[o][4], tx0, ty0)) {
        if (cnt < TILE_HIT_STRIDE - 1) out[1 + cnt] = (uint8_t)o;
        ++cnt;
      }
    out[0] = (uint8_t)(cnt <= TILE_HIT_STRIDE - 1 ? cnt : 255);
    mine += cnt;
  }
  // exclusive scan of the threads' counts, then one atomic per sample claims the sample's slice of the pair list
  s_scan[tid] = mine;
  __syncthreads();
  for (int d = 1; d < BIN_THREADS; d <<= 1) {
    const int v = tid >= d ? s_scan[tid - d] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  if (tid == BIN_THREADS - 1) {
    const int total = s_scan[BIN_THREADS - 1];
    const int base = atomicAdd(&a.pair_ctl[0], total);
    if (base + total > a.pair_cap) {  // more pairs than the mask buffer holds (the host sizes it from an upper bound, so this is a bug trap):
      a.pair_ctl[1] = 1;              // the raster and shade kernels skip the batch, and the host reports it after its next synchronisation
      if (a.pair_overflow) *a.pair_overflow = 1;
    }
    s_base = base;
  }
  __syncthreads();
  int off = s_base + s_scan[tid] - mine;
  const bool fits = s_base + s_scan[BIN_THREADS - 1] <= a.pair_cap;
  for (int t = tid + part * blockDim.x; t < n_tiles; t += blockDim.x * n_parts) {
    const int tx0 = (t % tiles_x) * TW, ty0 = (t / tiles_x) * TH;
    const int first = off;
    if (fits)
      for (int o = 0; o < n_obj; ++o)
        if (box_hits_tile(&s_box[o][0], tx0, ty0) || box_hits_tile(&s_box[o][4], tx0, ty0)) a.pair_list[off++] = make_int4(sample * 256 + o, t, s_shapes[o].x, s_shapes[o].y);
    a.tile_range[(size_t)sample * n_tiles + t] = make_int2(first, off - first);  // (count 0 if the list is full: the host sizes it from the boxes, so it never is)
  }
}

struct RasterSmem {  // shared memory of one block rasterising (object, tile) pairs
  int cover[RLAYER][RTH][TW];
  int area[RLAYER][RTH][TW];
  int carry[RLAYER][RTH];
  PairOutline out[NLAYER / 2];
  int seg_begin[NLAYER], seg_count[NLAYER];
  unsigned pairs[RASTER_ITEMS];
  int npairs;
  int next;
};

// One work unit (a pair, or a slice of RTH rows of it) by the whole block. kQueue: the block is persistent and claims its next
// unit from the queue a.pair_ctl[2] while it works on this one (pr_next / pe_next: the claimed unit and its record).
template <bool kDeform, bool kQueue>
__device__ __forceinline__ void raster_unit(const RenderArgs& a, RasterSmem& sm, int unit, int total, const int4 pe, int queue_base, int& pr_next,
                                            int4& pe_next) {
  const int W = a.W, H = a.H;
  const int tiles_x = (W + TW - 1) / TW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pr = unit / RSUB, slice = unit % RSUB;
  if (kQueue && tid == 0) sm.next = queue_base + atomicAdd(&a.pair_ctl[2], 1);
  const int tile = pe.y, shape_begin = pe.z;
  const int tx0 = (tile % tiles_x) * TW, ty0 = (tile / tiles_x) * TH + slice * RTH;
  const int x0 = tx0 + lane * 4;  // the warp sweeps tile rows warp, warp + RWARPS, ..
  const int n_shapes = pe.w & 0xFFFF, composite = pe.w >> 16;
  uint32_t uaa[RPW][2], una[RPW][2];
#pragma unroll
  for (int j = 0; j < RPW; ++j) uaa[j][0] = uaa[j][1] = una[j][0] = una[j][1] = 0u;
  if (a.pair_rows && tid >= RASTER_THREADS - 32 && lane < RTH) {
    // the object's span-interpolator rows over this tile (frame 1 = its texture under the inverse motion, DG.cpp:203-221),
    // one row per lane of the last warp (the (edge, row) items keep the first warps busy): the shade kernel's eight warps
    // pick them up instead of each working out its own
    const FlatObject& ob = a.objects[a.samples[pe.x >> 8].obj_begin + (pe.x & 255)];
    RowWarp rw;
    rw.init(ob.tex_inv, (double)(ty0 + lane), W);
    int4* r = a.pair_rows + (size_t)pr * PAIR_ROW_STRIDE;
    r[2 + (slice * RTH + lane) * 2] = make_int4(rw.dx.v1, rw.dx.lft, rw.dx.rem, rw.dy.v1);
    r[3 + (slice * RTH + lane) * 2] = make_int4(rw.dy.lft, rw.dy.rem, 0, 0);
    if (lane == 0 && slice == 0) {  // what the shade kernel needs of the object, in one record (instead of pair -> object -> texture table)
      const TexInfo ti = a.tex_info[ob.tex];
      r[0] = make_int4((int)(uint32_t)(ti.fg_base & 0xFFFFFFFFu), (int)(uint32_t)(ti.fg_base >> 32), ti.fg_pitch, pe.x & 255);
      r[1] = make_int4(ob.field, 0, 0, 0);
    }
  }
  for (int s0 = 0; s0 < n_shapes; s0 += RLAYER / 2) {
    const int ns = min(RLAYER / 2, n_shapes - s0);
    __syncthreads();  // the previous chunk (or pair) is done with the staging and the accumulators
    if (tid < ns) {   // outline tid of this chunk: which of its frames touch the tile, which layers they get
      const FlatShape& sh = a.shapes[shape_begin + s0 + tid];
      PairOutline po;
      const bool h0 = box_hits_rows(sh.bbox[0], tx0, ty0, RTH), h1 = box_hits_rows(sh.bbox[1], tx0, ty0, RTH);
      po.deform = (kDeform && h1) ? sh.deform : -1;  // a warped outline's frame-1 masks were materialised by the pre-pass
      const bool r1 = h1 && po.deform < 0;
      po.vbegin[0] = sh.vbegin[0]; po.vbegin[1] = sh.vbegin[1];
      po.vcount[0] = h0 ? sh.vcount[0] : 0; po.vcount[1] = r1 ? sh.vcount[1] : 0;
      po.layer[0] = h0 ? (signed char)(2 * tid) : (signed char)-1;
      po.layer[1] = r1 ? (signed char)(2 * tid + 1) : (signed char)-1;
      po.additive = sh.additive ? 1 : 0;
      sm.out[tid] = po;
      sm.seg_begin[2 * tid] = po.vbegin[0]; sm.seg_count[2 * tid] = po.vcount[0];
      sm.seg_begin[2 * tid + 1] = po.vbegin[1]; sm.seg_count[2 * tid + 1] = po.vcount[1];
    } else if (tid < NLAYER / 2) {
      sm.seg_count[2 * tid] = 0; sm.seg_count[2 * tid + 1] = 0;
    }
    for (int i = tid; i < 2 * ns * (RTH * TW / 4); i += RASTER_THREADS) {  // outline k owns layers 2k and 2k + 1
      reinterpret_cast<int4*>(&sm.cover[0][0][0])[i] = make_int4(0, 0, 0, 0);
      reinterpret_cast<int4*>(&sm.area[0][0][0])[i] = make_int4(0, 0, 0, 0);
    }
    if (tid < RLAYER * RTH) (&sm.carry[0][0])[tid] = 0;
    if (tid == 0) sm.npairs = 0;
    __syncthreads();
    if (kQueue && s0 == 0) {  // the claimed pair's record is in flight while this pair is rasterised
      pr_next = sm.next;
      if (pr_next < total) pe_next = a.pair_list[pr_next / RSUB];
    }
    {
      // (a) threads over edges: which tile rows does the edge cross? One work item per (edge, row).
      const int c0 = sm.seg_count[0], c1 = c0 + sm.seg_count[1], c2 = c1 + sm.seg_count[2], c3 = c2 + sm.seg_count[3];
      for (int e = tid; e < c3; e += RASTER_THREADS) {
        const int l = e < c0 ? 0 : (e < c1 ? 1 : (e < c2 ? 2 : 3));
        const int ei = e - (l == 0 ? 0 : (l == 1 ? c0 : (l == 2 ? c1 : c2)));
        const int n = sm.seg_count[l];
        const FlatVertex* v = a.verts + sm.seg_begin[l];
        const FlatVertex p = v[ei], q = v[ei + 1 == n ? 0 : ei + 1];
        int rlo, rhi;
        bool left;
        if (!tile_edge_rows(tx0, ty0, p.x, p.y, q.x, q.y, rlo, rhi, left, RTH)) continue;
        const int nrows = rhi - rlo + 1;
        const int base = left ? RASTER_ITEMS : atomicAdd(&sm.npairs, nrows);
        for (int k = 0; k < nrows; ++k) {
          if (base + k < RASTER_ITEMS) sm.pairs[base + k] = ((unsigned)l << 28) | ((unsigned)(rlo + k - ty0) << 24) | (unsigned)ei;
          else tile_edge_row<true>(&sm.cover[l][0][0], &sm.area[l][0][0], &sm.carry[l][0], tx0, ty0, rlo + k, p.x, p.y, q.x, q.y);  // cheap (left of the tile) or list full
        }
      }
    }
    __syncthreads();
    {
      // (b) threads over (edge, row) items: closed-form row segment -> cells
      const int np = min(sm.npairs, RASTER_ITEMS);
      // (items stay packed in the first warps: spread over all warps, lane * RTH + warp, the same few dozen divergent items
      // issue from eight half-empty warps instead of two -- measured 0.160 -> 0.171 ms)
      for (int i = tid; i < np; i += RASTER_THREADS) {
        const unsigned w = sm.pairs[i];
        const int l = (int)(w >> 28), r = ty0 + (int)((w >> 24) & 15u), ei = (int)(w & 0xFFFFFFu);
        const int n = sm.seg_count[l];
        const FlatVertex* v = a.verts + sm.seg_begin[l];
        const FlatVertex p = v[ei], q = v[ei + 1 == n ? 0 : ei + 1];
        tile_edge_row<true>(&sm.cover[l][0][0], &sm.area[l][0][0], &sm.carry[l][0], tx0, ty0, r, p.x, p.y, q.x, q.y);
      }
    }
    __syncthreads();

    for (int k = 0; k < ns; ++k)
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int row = warp + j * RWARPS, y = ty0 + row;
      const bool live = (y < H) && (x0 < W);
      uint32_t vaa[2] = {0, 0}, vna[2] = {0, 0};
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        const int l = sm.out[k].layer[f];
        if (l < 0) {
          if (kDeform && f == 1 && sm.out[k].deform >= 0 && live) warped_mask_words(a, sm.out[k].deform, x0, y, vaa[1], vna[1]);
          continue;
        }
        const int4 c4 = *reinterpret_cast<const int4*>(&sm.cover[l][row][lane * 4]);
        const int4 a4 = *reinterpret_cast<const int4*>(&sm.area[l][row][lane * 4]);
        const int carry = sm.carry[l][row];
        const bool cells = (c4.x | c4.y | c4.z | c4.w | a4.x | a4.y | a4.z | a4.w) != 0;
        if (!__any_sync(0xffffffffu, cells)) {
          // no outline crosses this row inside the tile: coverage is constant along it
          if (carry == 0) continue;
          const int cv = coverage_alpha(carry, 0);
          vaa[f] = graylut((unsigned)cv) * 0x01010101u;
          vna[f] = cv >= 128 ? 0xFFFFFFFFu : 0u;
          continue;
        }
        int c[4] = {c4.x, c4.y, c4.z, c4.w}, ar[4] = {a4.x, a4.y, a4.z, a4.w};
        c[1] += c[0]; c[2] += c[1]; c[3] += c[2];
        int tot = c[3];  // warp-level inclusive prefix sum over the lanes' cover totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int o = __shfl_up_sync(0xffffffffu, tot, d);
          if (lane >= d) tot += o;
        }
        const int base = tot - c[3] + carry;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cv = coverage_alpha(base + c[i], ar[i]);
          vaa[f] |= graylut((unsigned)cv) << (8 * i);        // gamma_none
          vna[f] |= (cv >= 128 ? 255u : 0u) << (8 * i);      // gamma_threshold(0.5), then graylut(255) = 255
        }
      }

      if (composite) {
        if (s0 + k == 0) { uaa[j][0] = uaa[j][1] = una[j][0] = una[j][1] = 0; }
        const bool add = sm.out[k].additive != 0;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          uaa[j][f] = comp4_lut(uaa[j][f], vaa[f], add, a.comp_lut);
          // non-AA masks stay in {0, 255} (the rules are closed on it) unless a warp field resampled them
          if (kDeform) una[j][f] = comp4_lut(una[j][f], vna[f], add, a.comp_lut);
          else una[j][f] = add ? (una[j][f] | vna[f]) : (una[j][f] & ~vna[f]);
        }
      } else {
        uaa[j][0] = vaa[0]; uaa[j][1] = vaa[1]; una[j][0] = vna[0]; una[j][1] = vna[1];
      }
    }
  }
  // the object's four masks over this tile: [AA 0, AA 1, non-AA 0, non-AA 1][tile row][lane], one word = four pixels
#pragma unroll
  for (int j = 0; j < RPW; ++j) {
    uint32_t* pm = a.pair_masks + (size_t)pr * (4 * TH * 32) + (slice * RTH + warp + j * RWARPS) * 32 + lane;
    pm[0 * TH * 32] = uaa[j][0]; pm[1 * TH * 32] = uaa[j][1]; pm[2 * TH * 32] = una[j][0]; pm[3 * TH * 32] = una[j][1];
  }
  if (kQueue && n_shapes <= 0) {  // (an object without outlines: no barrier has published the claim yet)
    __syncthreads();
    pr_next = sm.next;
    if (pr_next < total) pe_next = a.pair_list[pr_next / RSUB];
    __syncthreads();
  }
}

template <bool kDeform>
__global__ void __launch_bounds__(RASTER_THREADS, OFDG_RASTER_MIN_BLOCKS) raster_pairs_kernel(RenderArgs a) {
  __shared__ RasterSmem sm;
  if (a.pair_ctl[1]) return;
  const int total = a.pair_ctl[0] * RSUB;  // work units: RSUB slices per pair
  int4 pe_next = (int)blockIdx.x < total ? a.pair_list[blockIdx.x / RSUB] : make_int4(0, 0, 0, 0);
  // Units differ a lot in cost (1 to 7 outlines, a few to hundreds of edges): after its first unit a block claims the next
  // one from a queue (pair_ctl[2]) instead of striding over the list, so no block is left with a long tail of heavy units.
  int unit = blockIdx.x;
  while (unit < total) {
    const int4 pe = pe_next;
    int pr_next = total;
    raster_unit<kDeform, true>(a, sm, unit, total, pe, (int)gridDim.x, pr_next, pe_next);
    unit = pr_next;
  }
}

// (Blocks of 4 or 2 rows of a tile instead of 8 -- a block's slot is only free again when its slowest warp is done -- were
// measured: 2 % faster than the same code at 8 rows, but the extra index arithmetic costs the register allocation more than
// that at the 48-register cap: 0.132 -> 0.134 ms. One block per tile it stays.)
template <bool kDeform, bool kExtra>
__global__ void __launch_bounds__(RENDER_THREADS, kDeform ? OFDG_SHADE_DEFORM_MIN_BLOCKS : OFDG_SHADE_MIN_BLOCKS) shade_kernel(RenderArgs a) {
  if (a.pair_ctl[1]) return;
  const int W = a.W, H = a.H;
  const int tiles_x = (W + TW - 1) / TW;
  const int tx0 = (blockIdx.x % tiles_x) * TW, ty0 = (blockIdx.x / tiles_x) * TH;
  const int sample = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int y = ty0 + warp, x0 = tx0 + lane * 4;
  const bool live = (y < H) && (x0 < W);
  const FlatSample& smp = a.samples[sample];
  const size_t P = (size_t)W * H;
  const int obj_begin = smp.obj_begin;
  uint32_t col0[4], col1[4];
  uint32_t id0 = 0, id1 = 0;  // four pixels, one byte each: 0 = background, k+1 = k-th foreground object (k < 255)
  const unsigned full = 0xffffffffu;
  const int2 range = a.tile_range[(size_t)sample * gridDim.x + blockIdx.x];
#if OFDG_SHADE_STAGE > 0
  // The tile's pairs are consecutive in the pair list, so their masks (4 KB each) and records are two contiguous runs: one thread
  // starts bulk copies of the first OFDG_SHADE_STAGE pairs into shared memory now, and they land while the background is filtered --
  // the pair loop then finds everything on chip instead of paying a global round trip per pair.
  __shared__ __align__(128) uint32_t s_masks[OFDG_SHADE_STAGE][4 * TH * 32];
  __shared__ __align__(16) int4 s_rows[OFDG_SHADE_STAGE][PAIR_ROW_STRIDE];
  __shared__ __align__(8) unsigned long long s_bar;
  if (range.y > 0) {  // (block-uniform)
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint32_t n = (uint32_t)min(range.y, OFDG_SHADE_STAGE);
      mbar_expect_tx(&s_bar, n * (uint32_t)(4 * TH * 32 * 4 + PAIR_ROW_STRIDE * 16));
      bulk_g2s(&s_masks[0][0], a.pair_masks + (size_t)range.x * (4 * TH * 32), n * (uint32_t)(4 * TH * 32 * 4), &s_bar);
      bulk_g2s(&s_rows[0][0], a.pair_rows + (size_t)range.x * PAIR_ROW_STRIDE, n * (uint32_t)(PAIR_ROW_STRIDE * 16), &s_bar);
    }
  }
#endif

  // ---- background: masks are all 255 (DG.cpp:684-690); frame 0 = centre window of the prepared
  //      texture, frame 1 = that texture warped by I^-1*M*I on the 2W x 2H canvas (DG.cpp:665-682)
  {
    const uchar4* bg = a.bg + (size_t)sample * (4 * P);
    const int W2 = 2 * W, H2 = 2 * H;
    // (lanes outside the frame take part in the votes with neutral values)
    const int yy = live ? y : 0, xx = live ? x0 : 0;
    const uchar4* row = bg + (size_t)(yy + H / 2) * W2 + (xx + W / 2);
    if ((W & 7) == 0) {  // the lane's four pixels are one aligned 128-bit load
      const uint4 c4 = *reinterpret_cast<const uint4*>(row);
      col0[0] = c4.x & 0xFFFFFFu; col0[1] = c4.y & 0xFFFFFFu; col0[2] = c4.z & 0xFFFFFFu; col0[3] = c4.w & 0xFFFFFFu;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) col0[i] = ld_px(row + i) & 0xFFFFFFu;
    }
    const int4* br = a.bg_rows + ((size_t)sample * H + yy) * 2;
    const int4 r0 = br[0], r1 = br[1];
#if OFDG_SHADE_STAGE == 0
    if (range.y > 0) {  // the first pair's record and mask rows: on their way while the background is filtered
      const char* nm = reinterpret_cast<const char*>(a.pair_masks + (size_t)range.x * (4 * TH * 32) + warp * 32);
      const int4* prow = a.pair_rows + (size_t)range.x * PAIR_ROW_STRIDE;
      if (lane < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(nm + lane * (TH * 32 * 4)));
      else if (lane == 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow));
      else if (lane == 5) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow + 2 + warp * 2));
    }
#endif
    const int shift = pow2_shift(W2);
    SpanLane sl;
    sl.init(r0, r1, xx + W / 2, W2, shift < 0 ? 0 : shift);
    if (shift >= 0 && __all_sync(full, sl.inside(W2, H2))) {
#pragma unroll
      for (int i = 0; i < 4; ++i) col1[i] = bilinear_inside(reinterpret_cast<const uint32_t*>(bg), W2, sl.x_hr(i), sl.y_hr(i));
    } else {
#pragma unroll 1
      for (int i = 0; i < 4; ++i) col1[i] = bilinear_rgbx_call(bg, W2, W2, H2, W2, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, xx + i + W / 2);
    }
    if (live) {
      if (kDeform && smp.bg_field >= 0) {
        // background with a warp field: the warped 2W x 2H texture is resampled through the resized,
        // doubled inverse field before the centre crop (DG.cpp:670-681, 1194-1201)
        const int fw = W + 1, fh = H + 1;
        const float* ifl = a.fields + ((size_t)smp.bg_field * 2 + 1) * 2 * fw * fh;
        for (int i = 0; i < 4; ++i) {
          const int X = x0 + i + W / 2, Y = y + H / 2;
          const float sx = X + resized_field2(ifl, fw, fh, X, Y, a), sy = Y + resized_field2(ifl + (size_t)fw * fh, fw, fh, X, Y, a);
          col1[i] = warped_dirichlet_rgbx(bg, W2, W2, H2, smp.bg_tex_inv, sx, sy);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) col0[i] = col1[i] = 0;
    }
  }


  // ---- the tile's (object, tile) pairs in z-order: masks from the raster kernel, blits as in the fused kernel.
  //      Control flow around the votes is warp-uniform: lanes outside the frame carry empty masks instead of leaving.
  const int shift_fg = pow2_shift(W);
#if OFDG_SHADE_STAGE > 0
  uint32_t stage_parity = 0;
#endif
  for (int pr = range.x; pr < range.x + range.y; ++pr) {
    uint32_t uaa[2] = {0u, 0u}, una[2] = {0u, 0u};
#if OFDG_SHADE_STAGE > 0
    const int slot = (pr - range.x) % OFDG_SHADE_STAGE;
    if (slot == 0) {
      if (pr > range.x) {  // the next run of pairs: every warp is done with the staged ones
        __syncthreads();
        if (tid == 0) {
          const uint32_t n = (uint32_t)min(range.x + range.y - pr, OFDG_SHADE_STAGE);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&s_bar, n * (uint32_t)(4 * TH * 32 * 4 + PAIR_ROW_STRIDE * 16));
          bulk_g2s(&s_masks[0][0], a.pair_masks + (size_t)pr * (4 * TH * 32), n * (uint32_t)(4 * TH * 32 * 4), &s_bar);
          bulk_g2s(&s_rows[0][0], a.pair_rows + (size_t)pr * PAIR_ROW_STRIDE, n * (uint32_t)(PAIR_ROW_STRIDE * 16), &s_bar);
        }
      }
      mbar_wait(&s_bar, stage_parity);
      stage_parity ^= 1u;
    }
    const int4* prow = &s_rows[slot][0];
    if (live) {
      const uint32_t* pm = &s_masks[slot][warp * 32 + lane];
      uaa[0] = pm[0 * TH * 32]; uaa[1] = pm[1 * TH * 32]; una[0] = pm[2 * TH * 32]; una[1] = pm[3 * TH * 32];
    }
    const int4 hdr = prow[0];  // {foreground view base, pitch, object}
#else
    const int4* prow = a.pair_rows + (size_t)pr * PAIR_ROW_STRIDE;
    if (live) {
      const uint32_t* pm = a.pair_masks + (size_t)pr * (4 * TH * 32) + warp * 32 + lane;
      uaa[0] = pm[0 * TH * 32]; uaa[1] = pm[1 * TH * 32]; una[0] = pm[2 * TH * 32]; una[1] = pm[3 * TH * 32];
    }
    const int4 hdr = prow[0];  // {foreground view base, pitch, object}: in flight together with the masks
    if (pr + 1 < range.x + range.y) {  // the next pair's record and mask rows: on their way while this pair is blended
      const char* nm = reinterpret_cast<const char*>(a.pair_masks + (size_t)(pr + 1) * (4 * TH * 32) + warp * 32);
      if (lane < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(nm + lane * (TH * 32 * 4)));
      else if (lane == 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow + PAIR_ROW_STRIDE));
      else if (lane == 5) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow + PAIR_ROW_STRIDE + 2 + warp * 2));
    }
#endif
    const bool row_hit = __any_sync(full, (uaa[0] | uaa[1] | una[0] | una[1]) != 0u);
    if (!row_hit && !a.dbg_masks) continue;  // the row is clear of this object
    const int k = hdr.w;
    struct { unsigned long long fg_base; int fg_pitch; } ti;
    ti.fg_base = (unsigned long long)(uint32_t)hdr.x | ((unsigned long long)(uint32_t)hdr.y << 32);
    ti.fg_pitch = hdr.z;
    struct { int field; } ob;
    ob.field = kDeform ? prow[1].x : -1;
    // the object's masks are complete: ids from the non-AA masks, colour through the AA (or non-AA) masks
    if (a.dbg_masks && k < a.dbg_max_objs && live) {
      uint8_t* mb = a.dbg_masks + ((size_t)sample * a.dbg_max_objs + k) * 4 * P + (size_t)y * W + x0;
      *reinterpret_cast<uint32_t*>(mb + 0 * P) = uaa[0]; *reinterpret_cast<uint32_t*>(mb + 1 * P) = uaa[1];
      *reinterpret_cast<uint32_t*>(mb + 2 * P) = una[0]; *reinterpret_cast<uint32_t*>(mb + 3 * P) = una[1];
    }
    const uint32_t kk = (uint32_t)(k + 1) * 0x01010101u;
    const uint32_t e0 = __vcmpeq4(una[0], 0xFFFFFFFFu), e1 = __vcmpeq4(una[1], 0xFFFFFFFFu);
    id0 = (id0 & ~e0) | (kk & e0);
    id1 = (id1 & ~e1) | (kk & e1);
    const uint32_t m0w = a.use_aa ? uaa[0] : una[0], m1w = a.use_aa ? uaa[1] : una[1];
    const uchar4* tex = a.pool + ti.fg_base;  // the W x H foreground view (centre crop, DG.cpp:99-102 with defaults, or the resized copy)
    if (m0w) {
      const uchar4* trow = tex + (size_t)y * ti.fg_pitch + x0;  // identity warp == copy
#if OFDG_SHADE_VEC_FG
      if (((ti.fg_base | (unsigned long long)ti.fg_pitch) & 3ull) == 0ull) {  // the lane's four texels are one aligned 128-bit load
        const uint4 t4 = *reinterpret_cast<const uint4*>(trow);
        const uint32_t tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned m0 = (m0w >> (8 * i)) & 255u;
          if (m0) col0[i] = blend_rgbx(col0[i], tt[i] & 0xFFFFFFu, m0);
        }
      } else
#endif
      {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const unsigned m0 = (m0w >> (8 * i)) & 255u;
          if (m0) col0[i] = blend_rgbx(col0[i], ld_px(trow + i) & 0xFFFFFFu, m0);
        }
      }
    }
    if (!kDeform || ob.field < 0) {
      if (__any_sync(full, m1w != 0u)) {
        const int4 r0 = prow[2 + warp * 2], r1 = prow[3 + warp * 2];
        SpanLane sl;
        sl.init(r0, r1, live ? x0 : 0, W, shift_fg < 0 ? 0 : shift_fg);
        if (shift_fg >= 0 && __all_sync(full, m1w == 0u || sl.inside(W, H))) {
          if (m1w) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const unsigned m1 = (m1w >> (8 * i)) & 255u;
              if (m1) col1[i] = blend_rgbx(col1[i], bilinear_inside(reinterpret_cast<const uint32_t*>(tex), ti.fg_pitch, sl.x_hr(i), sl.y_hr(i)), m1);
            }
          }
        } else if (m1w) {
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const unsigned m1 = (m1w >> (8 * i)) & 255u;
            if (m1) col1[i] = blend_rgbx(col1[i], bilinear_rgbx_call(tex, ti.fg_pitch, W, H, W, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, x0 + i), m1);
          }
        }
      }
    } else if (kDeform && m1w) {
      // applyWarpFieldToTexture(getTransformedTexture(tex0, M), iflow) evaluated where the mask is set:
      // each of the 4 float-bilinear taps is itself one AGG span-bilinear pixel (DG.cpp:341-345)
      const double* tinv = a.objects[obj_begin + k].tex_inv;
      const int fw = W + 1, fh = H + 1;
      const float* ifl = a.fields + ((size_t)ob.field * 2 + 1) * 2 * fw * fh;
      for (int i = 0; i < 4; ++i) {
        const unsigned m1 = (m1w >> (8 * i)) & 255u;
        if (!m1) continue;
        const int x = x0 + i;
        const float sx = x + ifl[(size_t)y * fw + x], sy = y + ifl[(size_t)fw * fh + (size_t)y * fw + x];
        col1[i] = blend_rgbx(col1[i], warped_dirichlet_rgbx(tex, ti.fg_pitch, W, H, tinv, sx, sy), m1);
      }
    }
  }

  if (!live) return;

  // Forward flow of the lane's four pixels. The motion of the top-most object is loaded once and kept while the id stays the
  // same (it changes at object borders only); the row terms y * shx, y * sy are shared by the four pixels. Every product and
  // sum is the one getPointFlow forms, in its order: ((x * sx + y * shx) + tx) + W, minus the saved x (DG.cpp:390-401, 697-712).
  float fxv[4], fyv[4];
  {
    unsigned cur = 0xFFFFFFFFu;
    double m0 = 0, m1 = 0, m4 = 0, m5 = 0, t2 = 0, t3 = 0, pre_x = 0, pre_y = 0, ixb = 0, sxb = 0, syd = 0;
    const FlatObject* fo = nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned oid = (id0 >> (8 * i)) & 255u;
      if (oid != cur) {
        cur = oid;
        fo = oid ? a.objects + obj_begin + oid - 1 : nullptr;
        const double* m = oid ? fo->motion : smp.bg_motion;
        pre_x = oid ? 0.0 : (double)W; pre_y = oid ? 0.0 : (double)H;
        sxb = (double)(oid ? x0 : x0 + W / 2); syd = (double)(oid ? y : y + H / 2);  // (double)(float)(integer): exact
        ixb = sxb - pre_x;
        const double iy0 = syd - pre_y;
        m0 = m[0]; m1 = m[1]; m4 = m[4]; m5 = m[5];
        t2 = iy0 * m[2]; t3 = iy0 * m[3];
      }
      const double di = (double)i;
      const double ix0 = ixb + di, sx = sxb + di;  // small integers: exact
      double ix = ix0 * m0 + t2 + m4, iy = ix0 * m1 + t3 + m5;
      ix = ix + pre_x; iy = iy + pre_y;
      fxv[i] = (float)(ix - sx);
      fyv[i] = (float)(iy - syd);
      if (kDeform) {  // the forward field is added (DG.cpp:403-406, 714-717)
        const int fw = W + 1, fh = H + 1;
        if (oid == 0) {
          if (smp.bg_field >= 0 && ix >= 0 && ix < 2 * W && iy >= 0 && iy < 2 * H) {  // DG.cpp:714-717
            const float* fl = a.fields + ((size_t)smp.bg_field * 2 + 0) * 2 * fw * fh;
            auto at0 = [&](unsigned X, unsigned Y) { return resized_field2(fl, fw, fh, (int)X, (int)Y, a); };
            auto at1 = [&](unsigned X, unsigned Y) { return resized_field2(fl + (size_t)fw * fh, fw, fh, (int)X, (int)Y, a); };
            fxv[i] += neumann_f(at0, 2 * W, 2 * H, (float)ix, (float)iy);
            fyv[i] += neumann_f(at1, 2 * W, 2 * H, (float)ix, (float)iy);
          }
        } else if (fo->field >= 0 && ix >= 0 && ix < W && iy >= 0 && iy < H) {  // DG.cpp:403-406
          const float* fl = a.fields + ((size_t)fo->field * 2 + 0) * 2 * fw * fh;
          auto at0 = [&](unsigned X, unsigned Y) { return fl[(size_t)Y * fw + X]; };
          auto at1 = [&](unsigned X, unsigned Y) { return fl[(size_t)fw * fh + (size_t)Y * fw + X]; };
          fxv[i] += neumann_f(at0, fw, fh, (float)ix, (float)iy);
          fyv[i] += neumann_f(at1, fw, fh, (float)ix, (float)iy);
        }
      }
    }
  }


  // ---- write the three blobs (NCHW float): 8 planes x one 128-bit store per lane
  const size_t pix = (size_t)y * W + x0;
  float* of = a.flow + (size_t)sample * 2 * P + pix;
  if (a.img0) {
    float* o0 = a.img0 + (size_t)sample * 3 * P + pix;
    float* o1 = a.img1 + (size_t)sample * 3 * P + pix;
    if (smp.aug.enabled != 0) {  // this repository's own colour/noise augmentation (ofdg/augment.h); never set by the reference path
      store_augmented(&smp.aug, col0[0], col0[1], col0[2], col0[3], col1[0], col1[1], col1[2], col1[3], o0, o1, P, (uint32_t)pix);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#if OFDG_SHADE_I2F >= 1
        const float4 v0 = make_float4(byte_to_float_cvt(col0[0], c), byte_to_float_cvt(col0[1], c), byte_to_float_cvt(col0[2], c), byte_to_float_cvt(col0[3], c));
#else
        const float4 v0 = make_float4(byte_to_float(col0[0], c), byte_to_float(col0[1], c), byte_to_float(col0[2], c), byte_to_float(col0[3], c));
#endif
#if OFDG_SHADE_I2F >= 2
        const float4 v1 = make_float4(byte_to_float_cvt(col1[0], c), byte_to_float_cvt(col1[1], c), byte_to_float_cvt(col1[2], c), byte_to_float_cvt(col1[3], c));
#else
        const float4 v1 = make_float4(byte_to_float(col1[0], c), byte_to_float(col1[1], c), byte_to_float(col1[2], c), byte_to_float(col1[3], c));
#endif
        __stcs(reinterpret_cast<float4*>(o0 + c * P), v0);
        __stcs(reinterpret_cast<float4*>(o1 + c * P), v1);
      }
    }
  }
  __stcs(reinterpret_cast<float4*>(of), make_float4(fxv[0], fxv[1], fxv[2], fxv[3]));
  __stcs(reinterpret_cast<float4*>(of + P), make_float4(fyv[0], fyv[1], fyv[2], fyv[3]));

  if (kExtra) {
    if (a.flow_bw) {
  // ---- flow of the top-most object, f64 -> f32 (DG.cpp:388-401, 692-712). Forward: frame 0's ids through the motions;
      //      backward (extra top, computeFlowImage(inverse = true)): frame 1's ids through the inverse motions.
      auto point_flow = [&](unsigned oid, bool inverse, int i, float& fx, float& fy) {
        const float xf = (float)(x0 + i), yf = (float)y;
        // background: the point goes through I^-1 = T(-W,-H), M, I = T(W,H) (DG.cpp:697-712); objects: through M alone.
        // One code path: the translations are exact no-ops (+-0.0) for objects.
        const FlatObject* fo = oid ? a.objects + obj_begin + oid - 1 : nullptr;
        const double* m = oid ? (inverse ? fo->tex_inv : fo->motion) : (inverse ? smp.bg_motion_inv : smp.bg_motion);
        const double pre_x = oid ? 0.0 : (double)W, pre_y = oid ? 0.0 : (double)H;
        const float save_x = oid ? xf : xf + (float)(W / 2), save_y = oid ? yf : yf + (float)(H / 2);
        double ix = (double)save_x - pre_x, iy = (double)save_y - pre_y;
        const double tmp = ix;
        ix = tmp * m[0] + iy * m[2] + m[4];
        iy = tmp * m[1] + iy * m[3] + m[5];
        ix = ix + pre_x; iy = iy + pre_y;
        fx = (float)(ix - save_x);
        fy = (float)(iy - save_y);
        if (kDeform) {  // the forward field is added in both directions (DG.cpp:403-406, 714-717)
          const int fw = W + 1, fh = H + 1;
          if (oid == 0) {
            if (smp.bg_field >= 0 && ix >= 0 && ix < 2 * W && iy >= 0 && iy < 2 * H) {  // DG.cpp:714-717
              const float* fl = a.fields + ((size_t)smp.bg_field * 2 + 0) * 2 * fw * fh;
              auto at0 = [&](unsigned X, unsigned Y) { return resized_field2(fl, fw, fh, (int)X, (int)Y, a); };
              auto at1 = [&](unsigned X, unsigned Y) { return resized_field2(fl + (size_t)fw * fh, fw, fh, (int)X, (int)Y, a); };
              fx += neumann_f(at0, 2 * W, 2 * H, (float)ix, (float)iy);
              fy += neumann_f(at1, 2 * W, 2 * H, (float)ix, (float)iy);
            }
          } else if (fo->field >= 0 && ix >= 0 && ix < W && iy >= 0 && iy < H) {  // DG.cpp:403-406
            const float* fl = a.fields + ((size_t)fo->field * 2 + 0) * 2 * fw * fh;
            auto at0 = [&](unsigned X, unsigned Y) { return fl[(size_t)Y * fw + X]; };
            auto at1 = [&](unsigned X, unsigned Y) { return fl[(size_t)fw * fh + (size_t)Y * fw + X]; };
            fx += neumann_f(at0, fw, fh, (float)ix, (float)iy);
            fy += neumann_f(at1, fw, fh, (float)ix, (float)iy);
          }
        }
      };
      float bx[4], by[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) point_flow((id1 >> (8 * i)) & 255u, true, i, bx[i], by[i]);
      float* ob = a.flow_bw + (size_t)sample * 2 * P + pix;
      __stcs(reinterpret_cast<float4*>(ob), make_float4(bx[0], bx[1], bx[2], bx[3]));
      __stcs(reinterpret_cast<float4*>(ob + P), make_float4(by[0], by[1], by[2], by[3]));
    }
    if (a.top_id0 || a.top_id1) {
      float v0[4], v1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned b0 = (id0 >> (8 * i)) & 255u, b1 = (id1 >> (8 * i)) & 255u;
        v0[i] = (float)(b0 ? a.objects[obj_begin + b0 - 1].obj_id : 1);  // the background's ID is 1 (data_generation_layer.cpp:199)
        v1[i] = (float)(b1 ? a.objects[obj_begin + b1 - 1].obj_id : 1);
      }
      if (a.top_id0) __stcs(reinterpret_cast<float4*>(a.top_id0 + (size_t)sample * P + pix), make_float4(v0[0], v0[1], v0[2], v0[3]));
      if (a.top_id1) __stcs(reinterpret_cast<float4*>(a.top_id1 + (size_t)sample * P + pix), make_float4(v1[0], v1[1], v1[2], v1[3]));
    }
    if (a.ids8) {
      *reinterpret_cast<uint32_t*>(a.ids8 + (size_t)sample * 2 * P + pix) = id0;
      *reinterpret_cast<uint32_t*>(a.ids8 + (size_t)sample * 2 * P + P + pix) = id1;
    }
  }

  if (a.dbg_id0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned b0 = (id0 >> (8 * i)) & 255u, b1 = (id1 >> (8 * i)) & 255u;
      const unsigned o0id = b0 ? (unsigned)a.objects[obj_begin + b0 - 1].obj_id : 1u;
      const unsigned o1id = b1 ? (unsigned)a.objects[obj_begin + b1 - 1].obj_id : 1u;
      a.dbg_id0[(size_t)sample * P + pix + i] = o0id;
      if (a.dbg_id1) a.dbg_id1[(size_t)sample * P + pix + i] = o1id;
    }
  }
  if (a.frames8) {  // byte planes: channel c of the lane's four pixels is one 32-bit store
    uint8_t* fb = a.frames8 + (size_t)sample * 6 * P + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned sel = 0x40u + (unsigned)c * 0x11u;  // byte c of the first operand, byte c of the second
      const uint32_t w0 = __byte_perm(__byte_perm(col0[0], col0[1], sel), __byte_perm(col0[2], col0[3], sel), 0x5410u);
      const uint32_t w1 = __byte_perm(__byte_perm(col1[0], col1[1], sel), __byte_perm(col1[2], col1[3], sel), 0x5410u);
      __stcs(reinterpret_cast<uint32_t*>(fb + c * P), w0);
      __stcs(reinterpret_cast<uint32_t*>(fb + (3 + c) * P), w1);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// mode 9 pre-pass: frame-1 masks of warped outlines (MovingObjectBase::renderMasks, DG.cpp:370-386)
// ------------------------------------------------------------------------------------------------
// (a) rasterise the outline's frame-1 AA / non-AA masks over the whole frame, tile by tile
__global__ void __launch_bounds__(RENDER_THREADS) deform_raster_kernel(RenderArgs a) {
  __shared__ int s_cover[TH][TW];
  __shared__ int s_area[TH][TW];
  __shared__ int s_carry[TH];
  const int W = a.W, H = a.H;
  const size_t P = (size_t)W * H;
  const int tiles_x = (W + TW - 1) / TW;
  const int tx0 = (blockIdx.x % tiles_x) * TW, ty0 = (blockIdx.x / tiles_x) * TH;
  const int slot = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int y = ty0 + warp, x0 = tx0 + lane * 4;
  const bool live = (y < H) && (x0 < W);
  const FlatShape& sh = a.shapes[a.deform_shape[slot]];
  uint8_t* out = a.mask_raw + (size_t)slot * 2 * P + (size_t)y * W + x0;
  uint32_t paa = 0, pna = 0;
  if (!box_hits_tile(sh.raw1, tx0, ty0)) return;  // the warp pass reads nothing outside the outline's box (block-uniform)
  {
    for (int i = tid; i < TH * TW; i += RENDER_THREADS) { (&s_cover[0][0])[i] = 0; (&s_area[0][0])[i] = 0; }
    if (tid < TH) s_carry[tid] = 0;
    __syncthreads();
    const int n = sh.vcount[1];
    const FlatVertex* v = a.verts + sh.vbegin[1];
    for (int e = tid; e < n; e += RENDER_THREADS) {
      const FlatVertex p = v[e], q = v[e + 1 == n ? 0 : e + 1];
      tile_edge<true>(&s_cover[0][0], &s_area[0][0], &s_carry[0], tx0, ty0, p.x, p.y, q.x, q.y);
    }
    __syncthreads();
    const int4 c4 = *reinterpret_cast<const int4*>(&s_cover[warp][lane * 4]);
    const int4 a4 = *reinterpret_cast<const int4*>(&s_area[warp][lane * 4]);
    int c[4] = {c4.x, c4.y, c4.z, c4.w}, ar[4] = {a4.x, a4.y, a4.z, a4.w};
    c[1] += c[0]; c[2] += c[1]; c[3] += c[2];
    int tot = c[3];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, tot, d);
      if (lane >= d) tot += o;
    }
    const int base = tot - c[3] + s_carry[warp];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cv = coverage_alpha(base + c[i], ar[i]);
      paa |= graylut((unsigned)cv) << (8 * i);
      pna |= (cv >= 128 ? 255u : 0u) << (8 * i);
    }
  }
  if (live) {
    *reinterpret_cast<uint32_t*>(out) = paa;
    *reinterpret_cast<uint32_t*>(out + P) = pna;
  }
}
// (b) applyWarpFieldToTexture(mask, iflow): out(x,y) = trunc(bilinear_0(mask, (x,y) + iflow(x,y))), evaluated over the
// outline's warp region only (one thread = one four-pixel word of it; the raw mask is 0 outside the outline's own box, where
// pass (a) wrote nothing)
__global__ void deform_warp_kernel(RenderArgs a) {
  const int W = a.W, H = a.H;
  const size_t P = (size_t)W * H;
  const int slot = blockIdx.y >> 1, which = blockIdx.y & 1;
  const FlatShape& sh = a.shapes[a.deform_shape[slot]];
  const WarpRegion r = warp_region(sh.bbox[1], W, H);
  if (r.empty()) return;
  const int wpr = (r.x1 - r.x0 + 4) >> 2;  // words per region row
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int row = i / wpr;
  if (row > r.y1 - r.y0) return;
  const int y = r.y0 + row, xw = r.x0 + 4 * (i - row * wpr);
  const int bx0 = max(sh.raw1[0], 0), by0 = max(sh.raw1[1], 0), bx1 = min(sh.raw1[2], W - 1), by1 = min(sh.raw1[3], H - 1);
  const int fw = W + 1, fh = H + 1;
  const float* ifl = a.fields + ((size_t)a.deform_field[slot] * 2 + 1) * 2 * fw * fh;
  const uint8_t* src = a.mask_raw + ((size_t)slot * 2 + which) * P;
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = xw + k;
    if (x >= W) break;
    const float fx = x + ifl[(size_t)y * fw + x], fy = y + ifl[(size_t)fw * fh + (size_t)y * fw + x];
    int ix, iy;
    float dx, dy;
    if (!dirichlet_setup(fx, fy, ix, iy, dx, dy)) continue;
    auto tap = [&](int px, int py) -> float { return (px < bx0 || py < by0 || px > bx1 || py > by1) ? 0.f : (float)src[(size_t)py * W + px]; };
    const float v = cimg_lerp2(tap(ix, iy), tap(ix + 1, iy), tap(ix, iy + 1), tap(ix + 1, iy + 1), dx, dy);
    out |= (uint32_t)(unsigned char)v << (8 * k);
  }
  if (xw + 3 < W) *reinterpret_cast<uint32_t*>(a.mask_warp + ((size_t)slot * 2 + which) * P + (size_t)y * W + xw) = out;
  else for (int k = 0; xw + k < W; ++k) a.mask_warp[((size_t)slot * 2 + which) * P + (size_t)y * W + xw + k] = (uint8_t)(out >> (8 * k));
}

// ------------------------------------------------------------------------------------------------
// background texture preparation (CImg chain)
// ------------------------------------------------------------------------------------------------
// Linear-resize tables of CImg::get_resize (interpolation 3, growing axis): the source position is
// accumulated by repeated double additions, so a table is a sequential chain -- but it depends on the
// source length only. All tables (len = 2 .. n-1 -> n) are produced once at start-up, one thread each.
__global__ void resize_tables_kernel(int* pos_all, double* alpha_all, int n) {
  const int len = blockIdx.x * blockDim.x + threadIdx.x;
  if (len < 2 || len >= n) return;
  int* pos = pos_all + (size_t)len * n;
  double* alpha = alpha_all + (size_t)len * n;
  const double f = n > 1 ? (len - 1.0) / (n - 1) : 0;
  double curr = 0, old = 0;
  unsigned q = 0;
  for (int i = 0; i < n; ++i) {
    alpha[i] = curr - (unsigned int)curr;
    pos[i] = (int)q;
    old = curr;
    curr = fmin(len - 1.0, curr + f);
    q += (unsigned int)curr - (unsigned int)old;
  }
}

constexpr int PT = 32;        // prepared-texture tile edge
constexpr int PS = 46;        // source tile edge: ceil(32 * 1.3) + slack  (zoom >= 0.8 => crop <= 1.25 * 2W; larger ratios: bg_prep_general)
constexpr int PREP_THREADS = 256;

// cimg::mod(float x, float m) = (float)(dx - dm * floor(dx / dm)) in double. For 0 <= x < m the quotient's
// floor is 0 and the result is x itself; for -m <= x < 0 it is -1 and the result is (float)(dx + dm).
__device__ __forceinline__ float cimg_mod_f(float x, float m) {
  if (x >= 0.f && x < m) return x;
  const double dx = (double)x, dm = (double)m;
  if (x < 0.f && x >= -m) return (float)(dx + dm);
  return (float)(dx - dm * floor(dx / dm));
}

// pixel (x, y) of get_shift(sx, sy, 0, 0, mirror) of the pool texture
__device__ __forceinline__ uint32_t shifted_px(const uchar4* tex, int w, int h, int sx, int sy, int x, int y) {
  return ld_px(tex + (size_t)mirror(y - sy, h) * w + mirror(x - sx, w)) & 0xFFFFFFu;
}

// pixel (x, y) of rotate(angle, linear, mirror) of the shifted texture
__device__ __forceinline__ uint32_t rotated_px(const uchar4* tex, int w, int h, const BgPrep& p, int x, int y) {
  if (p.rot_identity) return shifted_px(tex, w, h, p.shift_x, p.shift_y, x, y);
  const float ww = 2.0f * w, hh = 2.0f * h;
  const float xc = x - p.rw2, yc = y - p.rh2;
  const float mx = cimg_mod_f(p.w2 + xc * p.ca + yc * p.sa, ww), my = cimg_mod_f(p.h2 - xc * p.sa + yc * p.ca, hh);
  const float fx = mx < w ? mx : ww - mx - 1, fy = my < h ? my : hh - my - 1;
  // _linear_atXY (Neumann)
  const float nfx = fx <= 0 ? 0 : (fx >= w - 1 ? (float)(w - 1) : fx), nfy = fy <= 0 ? 0 : (fy >= h - 1 ? (float)(h - 1) : fy);
  const unsigned int ix = (unsigned int)nfx, iy = (unsigned int)nfy;
  const float dx = nfx - ix, dy = nfy - iy;
  const unsigned int nx = dx > 0 ? ix + 1 : ix, ny = dy > 0 ? iy + 1 : iy;
  const uint32_t pcc = shifted_px(tex, w, h, p.shift_x, p.shift_y, ix, iy), pnc = shifted_px(tex, w, h, p.shift_x, p.shift_y, nx, iy),
                 pcn = shifted_px(tex, w, h, p.shift_x, p.shift_y, ix, ny), pnn = shifted_px(tex, w, h, p.shift_x, p.shift_y, nx, ny);
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float Icc = byte_to_float(pcc, c), Inc = byte_to_float(pnc, c), Icn = byte_to_float(pcn, c), Inn = byte_to_float(pnn, c);
    const float v = Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
    out |= ((uint32_t)(unsigned char)v) << (8 * c);
  }
  return out;
}

// Fast form of rotated_px for source tiles that provably stay clear of every boundary rule: the rotated sample positions
// lie strictly inside the texture (cimg::mod, the mirror fold and the Neumann clamp of get_rotate are identities) and the
// taps fall into one affine piece of get_shift's mirror boundary, column = cx0 + cs * X, row = ry0 + rs * Y.
struct TapMap { int cx0, cs, ry0, rs; };
__device__ __forceinline__ bool tap_axis(int lo, int hi, int shift, int n, int& c0, int& sg) {
  if (lo - shift >= 0 && hi - shift < n) { c0 = -shift; sg = 1; return true; }        // mirror(i, n) = i
  if (hi - shift < 0 && lo - shift >= -n) { c0 = shift - 1; sg = -1; return true; }   // mirror(i, n) = -i - 1
  return false;
}
__device__ __forceinline__ void rotated_pos(const BgPrep& p, int x, int y, float& fx, float& fy) {
  const float xc = x - p.rw2, yc = y - p.rh2;
  fx = p.w2 + xc * p.ca + yc * p.sa;
  fy = p.h2 - xc * p.sa + yc * p.ca;
}
// floor of 0 <= v < 2^22 without the conversion unit: adding 2^23 rounding down leaves floor(v) in the low mantissa bits,
// and subtracting 2^23 again gives it back as a float (both exact)
__device__ __forceinline__ unsigned floor_bits(float v, float& as_float) {
  const float t = __fadd_rd(v, 8388608.0f);
  as_float = t - 8388608.0f;
  return __float_as_uint(t) & 0x7FFFFFu;
}
__device__ __forceinline__ uint32_t rotated_px_fast(const uchar4* tex, int w, const BgPrep& p, const TapMap& m, float xf, float yf) {
  const float xc = xf - p.rw2, yc = yf - p.rh2;  // xf, yf: the pixel coordinates as floats (exact)
  const float fx = p.w2 + xc * p.ca + yc * p.sa, fy = p.h2 - xc * p.sa + yc * p.ca;
  float fix, fiy;
  const unsigned int ix = floor_bits(fx, fix), iy = floor_bits(fy, fiy);  // == (unsigned int)fx: fx, fy > 0
  const float dx = fx - fix, dy = fy - fiy;
  const uchar4* t00 = tex + (ptrdiff_t)(m.ry0 + m.rs * (int)iy) * w + (m.cx0 + m.cs * (int)ix);
  const ptrdiff_t ox = dx > 0 ? m.cs : 0, oy = dy > 0 ? (ptrdiff_t)m.rs * w : 0;
  const uint32_t pcc = ld_px(t00), pnc = ld_px(t00 + ox), pcn = ld_px(t00 + oy), pnn = ld_px(t00 + ox + oy);  // (the X byte is never selected)
  uint32_t t[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // CImg: v = Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc). The byte differences are small
    // integers, exact in float in any order -- so they are taken between the biased values 2^23 + byte directly (the
    // bias cancels exactly) and only the leading Icc is un-biased: three additions per channel less, same roundings.
    const float Mcc = byte_to_biased(pcc, c), Mnc = byte_to_biased(pnc, c), Mcn = byte_to_biased(pcn, c), Mnn = byte_to_biased(pnn, c);
    const float v = (Mcc - 8388608.0f) + dx * ((Mnc - Mcc) + dy * ((Mcc - Mcn) + (Mnn - Mnc))) + dy * (Mcn - Mcc);
    // (unsigned char)v for 0 <= v <= 255: truncate through the same mantissa trick (round towards zero)
    t[c] = __float_as_uint(__fadd_rz(fmaxf(v, 0.f), 8388608.0f));  // 0x4B0000vv (a rounding residue below zero also truncates to 0)
  }
  return __byte_perm(__byte_perm(t[0], t[1], 0x0040u), t[2], 0x5410u);  // {t0.b0, t1.b0, t2.b0, t2.b1 = 0}
}

// One resize pass (CImg get_resize interpolation 3; shrinking axes use the moving average)
// evaluated at output index t from a line of source pixels src[(s - s0) * stride].
// One output index of a resize pass (CImg get_resize interpolation 3; shrinking axes use the moving average)
// reduced to its taps, so that the index arithmetic is done once per column / row instead of once per pixel.
struct ResizeTaps {
  int first;      // first source index
  int count;      // 0: copy, 1..3: moving-average taps, -1: linear (first, first + 1 with weight alpha)
  unsigned w[3];  // moving average: overlap lengths, in accumulation order
  double alpha;
};
__device__ __forceinline__ ResizeTaps make_taps(int len, int n, int t, const int* pos, const double* alpha) {
  ResizeTaps r;
  r.alpha = 0.0; r.w[0] = r.w[1] = r.w[2] = 0u;
  if (len == n) { r.first = t; r.count = 0; }
  else if (len > n) {  // moving average over the exact rational overlap (at most 3 sources: len <= 1.3 n, checked on the host)
    const unsigned lo = (unsigned)t * (unsigned)len, hi = lo + (unsigned)len;  // < 2^31 for any supported size
    unsigned s = lo / (unsigned)n;
    r.first = (int)s;
    int k = 0;
    for (; s * (unsigned)n < hi && k < 3; ++s, ++k) {
      const unsigned b0 = max(s * (unsigned)n, lo), e0 = min((s + 1u) * (unsigned)n, hi);
      r.w[k] = e0 - b0;
    }
    r.count = k;
  } else { r.first = pos[t]; r.alpha = alpha[t]; r.count = -1; }
  return r;
}
// CImg's moving average accumulates byte * overlap in float and divides by the source length in float
// (get_resize, interpolation 2). Every partial sum is an integer below 2^24, so the float sums are exact; the
// quotient of two such integers is either an integer (exact in float) or at least 1/len away from the next one, far more
// than the float rounding error at values <= 255 -- so trunc(float(acc) / float(len)) == acc / len in integers.
// `magic` = floor(2^32 / len) + 1 turns that division into one multiply-high (exact for acc * len < 2^32).
// kMode: which of the three forms the pass takes is a property of the sample (crop length against prepared length), so the
// passes decide it once per block (0 copy, 1 moving average, 2 linear) instead of once per element (-1).
template <int kMode = -1>
__device__ __forceinline__ uint32_t apply_taps(const uint32_t* src, int stride, int s0, int len, unsigned magic, const ResizeTaps& r) {
  const uint32_t* p0 = src + (r.first - s0) * stride;
  if (kMode == 0 || (kMode < 0 && r.count == 0)) return p0[0];
  uint32_t t[3];  // the three channels' results, each below 256: packed by two byte permutes (t[2]'s byte 1 supplies the zero)
  if (kMode == 1 || (kMode < 0 && r.count > 0)) {
    unsigned acc[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < r.count) {
        const uint32_t p = p0[k * stride];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += __byte_perm(p, 0u, 0x4440u + (unsigned)c) * r.w[k];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = __umulhi(acc[c], magic);
    return __byte_perm(__byte_perm(t[0], t[1], 0x0040u), t[2], 0x5410u);
  }
  const uint32_t p1 = p0[0], p2 = r.first < len - 1 ? p0[stride] : p1;
  // byte -> double and double -> byte without the conversion unit: 2^52 + b has b in its low mantissa bits (exact), and
  // adding 2^52 with round-towards-zero leaves floor(v) there (0 <= v < 256)
  const double two52 = 4503599627370496.0, na = 1 - r.alpha;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double b1 = __hiloint2double(0x43300000, (int)__byte_perm(p1, 0u, 0x4440u + (unsigned)c)) - two52;
    const double b2 = __hiloint2double(0x43300000, (int)__byte_perm(p2, 0u, 0x4440u + (unsigned)c)) - two52;
    const double v = na * b1 + r.alpha * b2;
    t[c] = (uint32_t)__double2loint(__dadd_rz(v, two52));  // low word of 2^52 + floor(v): 0x000000vv
  }
  return __byte_perm(__byte_perm(t[0], t[1], 0x0040u), t[2], 0x5410u);
}

// fast path only (len <= 1.3 n): every product below 2^31
__device__ __forceinline__ void source_range32(int len, int n, int t0, int t1, const int* pos, int& s0, int& s1) {
  if (len == n) { s0 = t0; s1 = t1; }
  else if (len > n) { s0 = (int)(((unsigned)t0 * (unsigned)len) / (unsigned)n); s1 = (int)((((unsigned)(t1 + 1) * (unsigned)len) - 1u) / (unsigned)n); }
  else { s0 = pos[t0]; s1 = min(pos[t1] + 1, len - 1); }
}
__device__ __forceinline__ void source_range(int len, int n, int t0, int t1, const int* pos, int& s0, int& s1) {
  if (len == n) { s0 = t0; s1 = t1; }
  else if (len > n) { s0 = (int)(((long long)t0 * len) / n); s1 = (int)((((long long)(t1 + 1) * len) - 1) / n); }
  else { s0 = pos[t0]; s1 = min(pos[t1] + 1, len - 1); }
}

// One output index of a resize pass with any number of taps (general path; the fast path hoists its <= 3 taps).
__device__ __noinline__ uint32_t resize_general(const uint32_t* src, int stride, int s0, int len, int n, int t, const int* pos, const double* alpha) {
  if (len == n) return src[(t - s0) * stride];
  uint32_t out = 0;
  if (len > n) {  // moving average: float accumulation in source order, then one division (CImg interpolation 2)
    const unsigned lo = (unsigned)t * (unsigned)len, hi = lo + (unsigned)len;
    float acc[3] = {0.f, 0.f, 0.f};
    for (unsigned sidx = lo / (unsigned)n; sidx * (unsigned)n < hi; ++sidx) {
      const unsigned b0 = max(sidx * (unsigned)n, lo), e0 = min((sidx + 1u) * (unsigned)n, hi);
      const float wgt = (float)(e0 - b0);
      const uint32_t px = src[((int)sidx - s0) * stride];
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += byte_to_float(px, c) * wgt;
    }
    const float flen = (float)(unsigned int)len;
#pragma unroll
    for (int c = 0; c < 3; ++c) out |= ((uint32_t)(unsigned char)(acc[c] / flen)) << (8 * c);
    return out;
  }
  const int first = pos[t];
  const double al = alpha[t];
  const uint32_t p1 = src[(first - s0) * stride], p2 = first < len - 1 ? src[(first + 1 - s0) * stride] : p1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v = (1 - al) * (double)(int)((p1 >> (8 * c)) & 255u) + al * (double)(int)((p2 >> (8 * c)) & 255u);
    out |= ((uint32_t)(unsigned char)v) << (8 * c);
  }
  return out;
}

constexpr int PS_PITCH = PS;

// General path of the background preparation: the 32 x 32 output tile is cut into sub-tiles small enough for their
// source pixels to fit the shared staging area, whatever the resize ratio (textures smaller than 2W x 2H are resized
// whole, DataGenerator.cpp:103-107; rotated, they can be several times larger than 2W x 2H along one axis).
__device__ __noinline__ void bg_prep_general(const RenderArgs& a, const BgPrep& p, const uchar4* tex, int tex_w, int tex_h, int X0, int Y0, int X1, int Y1,
                                             const int* pos_x, const double* alpha_x, const int* pos_y, const double* alpha_y,
                                             uint32_t (*sA)[PS_PITCH], uint32_t (*sB)[PT], uchar4* out) {
  const int W2 = 2 * a.W, H2 = 2 * a.H;
  const int sub_w = p.crop_w > W2 ? max(1, min(PT, (int)(((long long)(PS - 2) * W2) / p.crop_w))) : PT;
  const int sub_h = p.crop_h > H2 ? max(1, min(PT, (int)(((long long)(PS - 2) * H2) / p.crop_h))) : PT;
  for (int ya = Y0; ya <= Y1; ya += sub_h)
    for (int xa = X0; xa <= X1; xa += sub_w) {
      const int xb = min(xa + sub_w - 1, X1), yb = min(ya + sub_h - 1, Y1);
      int cx0, cx1, cy0, cy1;
      source_range(p.crop_w, W2, xa, xb, pos_x, cx0, cx1);
      source_range(p.crop_h, H2, ya, yb, pos_y, cy0, cy1);
      const int cw = cx1 - cx0 + 1, ch = cy1 - cy0 + 1, tw = xb - xa + 1, th = yb - ya + 1;
      __syncthreads();  // the previous sub-tile is done with the staging area
      for (int i = threadIdx.x; i < cw * ch; i += PREP_THREADS) {
        const int lx = i % cw, ly = i / cw;
        sA[ly][lx] = rotated_px(tex, tex_w, tex_h, p, mirror(p.crop_x0 + cx0 + lx, p.rw), mirror(p.crop_y0 + cy0 + ly, p.rh));
      }
      __syncthreads();
      for (int i = threadIdx.x; i < tw * ch; i += PREP_THREADS) {
        const int lx = i % tw, ly = i / tw;
        sB[ly][lx] = resize_general(&sA[ly][0], 1, cx0, p.crop_w, W2, xa + lx, pos_x, alpha_x);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < tw * th; i += PREP_THREADS) {
        const int lx = i % tw, ly = i / tw;
        const uint32_t v = resize_general(&sB[0][lx], PT, cy0, p.crop_h, H2, ya + ly, pos_y, alpha_y);
        *reinterpret_cast<uint32_t*>(out + (size_t)(ya + ly) * W2 + xa + lx) = v;
      }
    }
}

#ifndef OFDG_PREP_MIN_BLOCKS
#define OFDG_PREP_MIN_BLOCKS 8  // measured: 4 / 5 / 6 / 8 blocks per SM -> 0.208 / 0.194 / 0.190 / 0.186 ms
#endif
struct PrepTileConst {
    int cx0, cy0, cw, ch, step_x, step_y, inv_cw, bx, by, inside, fast;
    unsigned magic_x, magic_y;
    TapMap tm;
};
struct PrepSmem {  // shared memory of one block preparing a 32 x 32 tile of a background
  uint32_t sA[PS][PS];  // rotated + cropped source pixels
  uint32_t sB[PS][PT];  // after the x pass
  ResizeTaps sTy[PT];   // taps of the tile's output rows
  PrepTileConst sC;
};

// One 32 x 32 tile (bx, by) of sample `sample`'s prepared background, by the whole block.
__device__ __forceinline__ void bg_prep_tile(const RenderArgs& a, PrepSmem& sm, int bx, int by, int sample) {
  uint32_t (*sA)[PS] = sm.sA;
  uint32_t (*sB)[PT] = sm.sB;
  ResizeTaps* sTy = sm.sTy;
  PrepTileConst& sC = sm.sC;
  typedef PrepTileConst TileConst;
  const BgPrep& p = a.samples[sample].prep;
  const int W2 = 2 * a.W, H2 = 2 * a.H;
  const int X0 = p.need[0] + bx * PT, Y0 = p.need[1] + by * PT;
  if (X0 > p.need[2] || Y0 > p.need[3]) return;
  const int X1 = min(X0 + PT - 1, p.need[2]), Y1 = min(Y0 + PT - 1, p.need[3]);
  // `need` is the bounding box of what the renderer reads: the centre W x H window (frame 0) and the footprint of the frame-1
  // warp, a parallelogram. A tile in a corner of that box that touches neither is skipped: its corners (three pixels of margin
  // for the filter taps and roundings), taken back through the forward transform, must bound a rectangle that meets the output
  // window. Not applied when the footprint left the canvas (reflect wrap: `need` then spans a whole axis), to backgrounds
  // with a warp field (they read anywhere) and to the parity instrumentation (a.flow == nullptr: the whole box is compared).
  if (a.flow && a.samples[sample].bg_field < 0 && !(p.need[0] == 0 && p.need[2] == W2 - 1) && !(p.need[1] == 0 && p.need[3] == H2 - 1)) {
    const bool frame0 = X0 <= a.W / 2 + a.W - 1 && X1 >= a.W / 2 && Y0 <= a.H / 2 + a.H - 1 && Y1 >= a.H / 2;
    if (!frame0) {
      const double* m = a.samples[sample].bg_tex_inv;  // prepared = m * output
      const double det = m[0] * m[3] - m[1] * m[2];
      double ux0 = 1e300, ux1 = -1e300, uy0 = 1e300, uy1 = -1e300;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double px = ((c & 1) ? X1 + 4 : X0 - 4) - m[4], py = ((c & 2) ? Y1 + 4 : Y0 - 4) - m[5];
        const double ox = (px * m[3] - py * m[2]) / det, oy = (py * m[0] - px * m[1]) / det;
        ux0 = fmin(ux0, ox); ux1 = fmax(ux1, ox); uy0 = fmin(uy0, oy); uy1 = fmax(uy1, oy);
      }
      // the output pixels whose centres (+0.5) the span interpolator maps: columns W/2 .. W/2 + W, rows H/2 .. H/2 + H
      if (ux1 < a.W / 2 - 1.0 || ux0 > a.W / 2 + a.W + 2.0 || uy1 < a.H / 2 - 1.0 || uy0 > a.H / 2 + a.H + 2.0) return;
    }
  }
  const int* pos_x = a.pos_x + (size_t)min(p.crop_w, W2 - 1) * W2;  // rows >= n are never read (no table needed when shrinking)
  const double* alpha_x = a.alpha_x + (size_t)min(p.crop_w, W2 - 1) * W2;
  const int* pos_y = a.pos_y + (size_t)min(p.crop_h, H2 - 1) * H2;
  const double* alpha_y = a.alpha_y + (size_t)min(p.crop_h, H2 - 1) * H2;
  const TexInfo ti = a.tex_info[p.tex];
  const uchar4* tex = a.pool + ti.off;
  uchar4* out = a.bg + (size_t)sample * W2 * H2;
  if (p.general) {
    bg_prep_general(a, p, tex, ti.w, ti.h, X0, Y0, X1, Y1, pos_x, alpha_x, pos_y, alpha_y, sA, sB, out);
    return;
  }
  // Per-tile constants (source ranges, multiply-high constants, the walk's steps, the fast-path test): a dozen integer
  // divisions and four corner evaluations that are the same for all 256 threads -- warp 0 works them out, the others
  // pick them up from shared memory (the kernel is issue-bound: seven warps' worth of redundant instructions saved).
  const int tw = X1 - X0 + 1, th = Y1 - Y0 + 1;
  if (threadIdx.x < 32) {
    TileConst c;
    int cx1, cy1;
    source_range32(p.crop_w, W2, X0, X1, pos_x, c.cx0, cx1);
    source_range32(p.crop_h, H2, Y0, Y1, pos_y, c.cy0, cy1);
    c.cw = cx1 - c.cx0 + 1; c.ch = cy1 - c.cy0 + 1;  // <= PS by construction
    c.magic_x = 0xFFFFFFFFu / (unsigned)p.crop_w + 1u; c.magic_y = 0xFFFFFFFFu / (unsigned)p.crop_h + 1u;
    c.step_y = PREP_THREADS / c.cw; c.step_x = PREP_THREADS % c.cw;
    c.inv_cw = 65536 / c.cw + 1;  // t / cw == (t * inv_cw) >> 16 for t < 256, cw <= 46
    c.bx = p.crop_x0 + c.cx0; c.by = p.crop_y0 + c.cy0;
    c.inside = c.bx >= 0 && c.by >= 0 && c.bx + c.cw <= p.rw && c.by + c.ch <= p.rh;  // the crop's mirror boundary is not in play for this tile
    c.fast = 0;
    c.tm = TapMap{0, 0, 0, 0};
    if (c.inside && !p.rot_identity) {  // the sample positions are affine in (x, y): their extremes sit at the tile's corners
      float x0f, y0f, x1f, y1f, x2f, y2f, x3f, y3f;
      rotated_pos(p, c.bx, c.by, x0f, y0f); rotated_pos(p, c.bx + c.cw - 1, c.by, x1f, y1f);
      rotated_pos(p, c.bx, c.by + c.ch - 1, x2f, y2f); rotated_pos(p, c.bx + c.cw - 1, c.by + c.ch - 1, x3f, y3f);
      const float xl = fminf(fminf(x0f, x1f), fminf(x2f, x3f)), xh = fmaxf(fmaxf(x0f, x1f), fmaxf(x2f, x3f));
      const float yl = fminf(fminf(y0f, y1f), fminf(y2f, y3f)), yh = fmaxf(fmaxf(y0f, y1f), fmaxf(y2f, y3f));
      if (xl >= 2.f && xh <= (float)(ti.w - 3) && yl >= 2.f && yh <= (float)(ti.h - 3))  // one texel of slack for rounding inside the tile
        c.fast = tap_axis((int)xl - 1, (int)xh + 2, p.shift_x, ti.w, c.tm.cx0, c.tm.cs) && tap_axis((int)yl - 1, (int)yh + 2, p.shift_y, ti.h, c.tm.ry0, c.tm.rs);
    }
    if (threadIdx.x == 0) sC = c;
  }
  if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + th)   // taps of the tile's output rows (warp 1; visible after the barriers below)
    sTy[threadIdx.x - 32] = make_taps(p.crop_h, H2, Y0 + (int)threadIdx.x - 32, pos_y, alpha_y);
  __syncthreads();
  const int cx0 = sC.cx0, cy0 = sC.cy0, cw = sC.cw, ch = sC.ch;
  const unsigned magic_x = sC.magic_x, magic_y = sC.magic_y;
  // A: crop(x0, y0, .., mirror) of the rotated image
  const int lane_x = threadIdx.x & 31, lane_y = threadIdx.x >> 5;
  {  // all lanes busy: walk the cw x ch source tile linearly, stepping (x, y) by 256 items without a divide per item
    const int step_y = sC.step_y, step_x = sC.step_x;
    int ly = ((int)threadIdx.x * sC.inv_cw) >> 16, lx = (int)threadIdx.x - ly * cw;
    const int bx = sC.bx, by = sC.by;
    const bool inside = sC.inside != 0, fast = sC.fast != 0;
    const TapMap tm = sC.tm;
    if (fast) {
      // The hot loop of the kernel (about 40 % of its instructions), written for the instruction count: everything that does not
      // change stays in registers, texels are addressed by a 32-bit index from the texture's base, the staging row is a running
      // shared-memory address, and the 2^23 bias the byte -> float permutes insert lives in a register so that their selectors
      // stay immediates. Arithmetic and roundings are those of rotated_px_fast (CImg's _linear_atXY formula).
      float xf = (float)(bx + lx), yf = (float)(by + ly);  // float twins of the walk (small integers: exact)
      const float fstep_x = (float)step_x, fstep_y = (float)step_y, fcw = (float)cw;
      const float ca = p.ca, sa = p.sa, w2 = p.w2, h2 = p.h2, rw2 = p.rw2, rh2 = p.rh2;
      const uint32_t* tex32 = reinterpret_cast<const uint32_t*>(tex);
      const int cs = tm.cs, rsw = tm.rs * ti.w, idx0 = tm.ry0 * ti.w + tm.cx0;  // texel index = idx0 + rsw * iy + cs * ix
      const uint32_t bias = a.float_bias;  // 0x4B000000
      uint32_t saddr = smem_u32(&sA[0][0]) + (uint32_t)(ly * PS + lx) * 4u;
      const uint32_t sstep = (uint32_t)(step_y * PS + step_x) * 4u, swrap = (uint32_t)(PS - cw) * 4u;
      while (ly < ch) {
        const float xc = xf - rw2, yc = yf - rh2;
        const float fx = w2 + xc * ca + yc * sa, fy = h2 - xc * sa + yc * ca;
        float fix, fiy;
        const unsigned int ix = floor_bits(fx, fix), iy = floor_bits(fy, fiy);  // == (unsigned int)fx: fx, fy > 0
        const float dx = fx - fix, dy = fy - fiy;
        const int i00 = idx0 + rsw * (int)iy + cs * (int)ix;
        const int ox = dx > 0 ? cs : 0, oy = dy > 0 ? rsw : 0;
        const uint32_t pcc = tex32[i00], pnc = tex32[i00 + ox], pcn = tex32[i00 + oy], pnn = tex32[i00 + ox + oy];
        uint32_t t[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float Mcc = __uint_as_float(__byte_perm(pcc, bias, 0x7440u + c)), Mnc = __uint_as_float(__byte_perm(pnc, bias, 0x7440u + c)),
                      Mcn = __uint_as_float(__byte_perm(pcn, bias, 0x7440u + c)), Mnn = __uint_as_float(__byte_perm(pnn, bias, 0x7440u + c));
          const float v = (Mcc - 8388608.0f) + dx * ((Mnc - Mcc) + dy * ((Mcc - Mcn) + (Mnn - Mnc))) + dy * (Mcn - Mcc);
          t[c] = __float_as_uint(__fadd_rz(fmaxf(v, 0.f), 8388608.0f));
        }
        const uint32_t px = __byte_perm(__byte_perm(t[0], t[1], 0x0040u), t[2], 0x5410u);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(px) : "memory");
        lx += step_x; ly += step_y; xf += fstep_x; yf += fstep_y; saddr += sstep;
        if (lx >= cw) { lx -= cw; ++ly; xf -= fcw; yf += 1.f; saddr += swrap; }
      }
    } else {
      while (ly < ch) {
        const int rx = inside ? bx + lx : mirror(bx + lx, p.rw), ry = inside ? by + ly : mirror(by + ly, p.rh);
        sA[ly][lx] = rotated_px(tex, ti.w, ti.h, p, rx, ry);
        lx += step_x; ly += step_y;
        if (lx >= cw) { lx -= cw; ++ly; }
      }
    }
  }
  __syncthreads();
  // B: resize along x (one column per lane: its taps are computed once)
  if (lane_x < tw) {
    const ResizeTaps tx = make_taps(p.crop_w, W2, X0 + lane_x, pos_x, alpha_x);
    if (p.crop_w == W2) {
      for (int ly = lane_y; ly < ch; ly += PREP_THREADS / 32) sB[ly][lane_x] = apply_taps<0>(&sA[ly][0], 1, cx0, p.crop_w, magic_x, tx);
    } else if (p.crop_w > W2) {
      for (int ly = lane_y; ly < ch; ly += PREP_THREADS / 32) sB[ly][lane_x] = apply_taps<1>(&sA[ly][0], 1, cx0, p.crop_w, magic_x, tx);
    } else {
      for (int ly = lane_y; ly < ch; ly += PREP_THREADS / 32) sB[ly][lane_x] = apply_taps<2>(&sA[ly][0], 1, cx0, p.crop_w, magic_x, tx);
    }
  }
  __syncthreads();
  // P: resize along y
  if (lane_x < tw) {
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + (size_t)Y0 * W2 + X0 + lane_x);
    if (p.crop_h == H2) {
      for (int ly = lane_y; ly < th; ly += PREP_THREADS / 32) orow[(size_t)ly * W2] = apply_taps<0>(&sB[0][lane_x], PT, cy0, p.crop_h, magic_y, sTy[ly]);
    } else if (p.crop_h > H2) {
      for (int ly = lane_y; ly < th; ly += PREP_THREADS / 32) orow[(size_t)ly * W2] = apply_taps<1>(&sB[0][lane_x], PT, cy0, p.crop_h, magic_y, sTy[ly]);
    } else {
      for (int ly = lane_y; ly < th; ly += PREP_THREADS / 32) orow[(size_t)ly * W2] = apply_taps<2>(&sB[0][lane_x], PT, cy0, p.crop_h, magic_y, sTy[ly]);
    }
  }
}

__global__ void __launch_bounds__(PREP_THREADS, OFDG_PREP_MIN_BLOCKS) bg_prep_kernel(RenderArgs a) {
  __shared__ PrepSmem sm;
  bg_prep_tile(a, sm, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

// Foreground view of a texture smaller than W x H: the whole texture resized (getRandomizedCrop's else branch with
// its default arguments, DataGenerator.cpp:103-107; shift 0 and angle 0 are copies). One-time, at upload.
__global__ void resize_table_kernel(int* pos, double* alpha, int len, int n) {  // CImg linear-resize table len -> n (len < n)
  if (threadIdx.x || blockIdx.x) return;
  const double f = n > 1 ? (len - 1.0) / (n - 1) : 0;
  double curr = 0, old = 0;
  unsigned q = 0;
  for (int i = 0; i < n; ++i) {
    alpha[i] = curr - (unsigned int)curr;
    pos[i] = (int)q;
    old = curr;
    curr = fmin(len - 1.0, curr + f);
    q += (unsigned int)curr - (unsigned int)old;
  }
}
__global__ void fg_resize_x_kernel(const uchar4* tex, int w, int h, int W, uint32_t* tmp, const int* pos, const double* alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * h) return;
  const int X = i % W, sy = i / W;
  tmp[i] = resize_general(reinterpret_cast<const uint32_t*>(tex) + (size_t)sy * w, 1, 0, w, W, X, pos, alpha) & 0xFFFFFFu;
}
__global__ void fg_resize_y_kernel(const uint32_t* tmp, int h, int W, int H, uchar4* out, const int* pos, const double* alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * H) return;
  const int X = i % W, Y = i / W;
  reinterpret_cast<uint32_t*>(out)[i] = resize_general(tmp + X, W, 0, h, H, Y, pos, alpha);
}

// ------------------------------------------------------------------------------------------------
// texture pool helpers
// ------------------------------------------------------------------------------------------------
__global__ void planar_to_rgbx_kernel(const uint8_t* planar, uchar4* out, size_t n_px, size_t plane) {
  // planar: [tex][3][plane]; out: [tex][plane]
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x) {
    const size_t t = i / plane, r = i % plane;
    const uint8_t* src = planar + t * 3 * plane + r;
    out[i] = make_uchar4(src[0], src[plane], src[2 * plane], 0);
  }
}
__global__ void rgbx_to_planar_kernel(const uchar4* in, uint8_t* planar, size_t plane) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 v = in[i];
    planar[i] = v.x; planar[plane + i] = v.y; planar[2 * plane + i] = v.z;
  }
}

__device__ __forceinline__ uint32_t synth_hash8(uint64_t key, uint32_t gx, uint32_t gy, uint32_t o) {
  uint64_t z = key + (uint64_t)gx * 0xBF58476D1CE4E5B9ull + (uint64_t)gy * 0x94D049BB133111EBull + (uint64_t)o * 0xD6E8FEB86659FD93ull;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z & 255u);
}
// Integer-only procedural texture: 4 octaves of value noise + a triangle-wave stripe + a checker.
// Mirrors ofdg_b200.synth_textures() (numpy) bit for bit.
__global__ void synth_textures_kernel(uchar4* out, int n, int w, int h, uint64_t seed, int first_index) {
  const size_t plane = (size_t)w * h, total = plane * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t t = (uint32_t)(i / plane) + (uint32_t)first_index;
    const uint32_t r = (uint32_t)(i % plane), x = r % w, y = r / w;
    uint32_t px[3];
    const uint64_t tkey = seed ^ ((uint64_t)(t + 1) * 0x9E3779B97F4A7C15ull);
    const uint32_t sax = synth_hash8(tkey, 1, 2, 77) % 17, say = synth_hash8(tkey, 3, 4, 77) % 17;  // stripe direction
    const uint32_t phase = (x * sax + y * say) & 255u;
    const uint32_t tri = (phase < 128u ? phase : 255u - phase) * 2u;  // 0..254
    const uint32_t checker = ((x >> 5) ^ (y >> 5)) & 1u;
    for (uint32_t c = 0; c < 3; ++c) {
      const uint64_t key = tkey + (uint64_t)(c + 1) * 0xA24BAED4963EE407ull;
      uint32_t acc = 0;
      for (uint32_t o = 0; o < 4; ++o) {
        const uint32_t cell = 64u >> o;
        const uint32_t gx = x / cell, gy = y / cell, fx = (x % cell) * 256u / cell, fy = (y % cell) * 256u / cell;
        const uint32_t v00 = synth_hash8(key, gx, gy, o), v10 = synth_hash8(key, gx + 1, gy, o),
                       v01 = synth_hash8(key, gx, gy + 1, o), v11 = synth_hash8(key, gx + 1, gy + 1, o);
        const uint32_t top = v00 * (256u - fx) + v10 * fx, bot = v01 * (256u - fx) + v11 * fx;
        const uint32_t val = (top * (256u - fy) + bot * fy) >> 16;
        acc += val << (3u - o);
      }
      uint32_t v = (acc / 15u) * 3u / 4u + tri / 4u + (checker ? 24u : 0u) + c * 5u;
      px[c] = v > 255u ? 255u : v;
    }
    out[i] = make_uchar4((unsigned char)px[0], (unsigned char)px[1], (unsigned char)px[2], 0);
  }
}

// Exhaustive table of the composite-mask rules as the render kernel evaluates them (parity check)
__global__ void composite_lut_kernel(uint8_t* add_lut, uint8_t* sub_lut) {
  __shared__ float q[256];
  q[threadIdx.x] = (float)threadIdx.x / 255.f;
  __syncthreads();
  const unsigned u = blockIdx.x, v = threadIdx.x;
  add_lut[u * 256 + v] = (uint8_t)comp_add(u, v, q);
  sub_lut[u * 256 + v] = (uint8_t)comp_sub(u, v, q);
}

}  // namespace

namespace {
__global__ void scene_upload_kernel(UploadSegments u) {
  const int seg = blockIdx.y;
  const uint4* src = static_cast<const uint4*>(u.src[seg]);
  uint4* dst = static_cast<uint4*>(u.dst[seg]);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < u.n16[seg]; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace

int launch_scene_upload(const UploadSegments& u, cudaStream_t s) {
  if (u.n <= 0) return 0;
  scene_upload_kernel<<<dim3(24, u.n), 256, 0, s>>>(u);
  return 1;
}

void launch_composite_luts(uint8_t* add_lut, uint8_t* sub_lut, cudaStream_t s) {
  composite_lut_kernel<<<256, 256, 0, s>>>(add_lut, sub_lut);
}

int launch_fg_resize(const uchar4* tex, int w, int h, int W, int H, uchar4* out, uint32_t* tmp, int* pos, double* alpha, cudaStream_t s) {
  // tmp: W x h pixels; pos / alpha: max(W, H) entries each (reused by the two passes, stream-ordered)
  int launches = 0;
  if (w < W) { resize_table_kernel<<<1, 1, 0, s>>>(pos, alpha, w, W); ++launches; }
  fg_resize_x_kernel<<<(W * h + 255) / 256, 256, 0, s>>>(tex, w, h, W, tmp, pos, alpha);
  if (h < H) { resize_table_kernel<<<1, 1, 0, s>>>(pos, alpha, h, H); ++launches; }
  fg_resize_y_kernel<<<(W * H + 255) / 256, 256, 0, s>>>(tmp, h, W, H, out, pos, alpha);
  return launches + 2;
}

void launch_resize_tables(int* pos, double* alpha, int n, cudaStream_t s) {
  resize_tables_kernel<<<(n + 63) / 64, 64, 0, s>>>(pos, alpha, n);
}

size_t tile_hits_bytes(int batch, int W, int H) {
  return (size_t)batch * ((W + TW - 1) / TW) * ((H + TH - 1) / TH) * TILE_HIT_STRIDE;
}
int launch_bin(const RenderArgs& a, cudaStream_t s) {
  bin_kernel<<<a.batch, 192, 0, s>>>(a);
  return 1;
}

int launch_background_prep(const RenderArgs& a, cudaStream_t s) {
  // blocks are placed relative to each sample's needed region (BgPrep::need): no block beyond the largest one has work
  const int pw = a.prep_w > 0 ? min(a.prep_w, 2 * a.W) : 2 * a.W, ph = a.prep_h > 0 ? min(a.prep_h, 2 * a.H) : 2 * a.H;
  dim3 grid((pw + PT - 1) / PT, (ph + PT - 1) / PT, a.batch);
  bg_prep_kernel<<<grid, PREP_THREADS, 0, s>>>(a);
  return 1;
}

int launch_deform_prepass(const RenderArgs& a, cudaStream_t s) {
  if (a.n_deform <= 0) return 0;
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH;
  deform_raster_kernel<<<dim3(tiles_x * tiles_y, a.n_deform), RENDER_THREADS, 0, s>>>(a);
  const size_t P = (size_t)a.W * a.H;
  deform_warp_kernel<<<dim3((unsigned)((P / 4 + 255) / 256 + 1), 2 * a.n_deform), 256, 0, s>>>(a);  // (blocks beyond an outline's region leave at once)
  return 2;
}

namespace {
// Occlusion top (this repository's definition, include/ofdg/ofdg.h): frame 0's pixel p, carried by its forward flow
// to the nearest frame-1 pixel q, is occluded when q lies outside the frame or shows another object there.
__global__ void occlusion_kernel(RenderArgs a) {
  const size_t P = (size_t)a.W * a.H;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int sample = blockIdx.y;
  if (i >= P) return;
  const int x = (int)(i % a.W), y = (int)(i / a.W);
  const float* fl = a.flow + (size_t)sample * 2 * P;
  const uint8_t* ids = a.ids8 + (size_t)sample * 2 * P;
  const float tx = (float)x + fl[i], ty = (float)y + fl[P + i];
  float occ = 1.f;
  if (tx >= -0.5f && tx < (float)a.W - 0.5f && ty >= -0.5f && ty < (float)a.H - 0.5f) {  // false for NaN
    const int qx = (int)floorf(tx + 0.5f), qy = (int)floorf(ty + 0.5f);
    if (ids[P + (size_t)qy * a.W + qx] == ids[i]) occ = 0.f;
  }
  a.occlusion[(size_t)sample * P + i] = occ;
}
}  // namespace

int launch_render(const RenderArgs& a, cudaStream_t s) {
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH;
  dim3 grid(tiles_x * tiles_y, a.batch);
  const bool extra = a.flow_bw || a.top_id0 || a.top_id1 || a.ids8;
  if (a.n_fields > 0) {
    if (extra) render_kernel<true, true><<<grid, RENDER_THREADS, 0, s>>>(a);
    else render_kernel<true, false><<<grid, RENDER_THREADS, 0, s>>>(a);
  } else {
    if (extra) render_kernel<false, true><<<grid, RENDER_THREADS, 0, s>>>(a);
    else render_kernel<false, false><<<grid, RENDER_THREADS, 0, s>>>(a);
  }
  if (!a.occlusion) return 1;
  const size_t P = (size_t)a.W * a.H;
  occlusion_kernel<<<dim3((unsigned)((P + 255) / 256), a.batch), 256, 0, s>>>(a);
  return 2;
}

size_t pair_mask_bytes_per_pair() { return (size_t)4 * TH * 32 * sizeof(uint32_t); }
size_t pair_row_bytes_per_pair() { return (size_t)PAIR_ROW_STRIDE * sizeof(int4); }

int launch_bin_pairs(const RenderArgs& a, cudaStream_t s) {
  cudaMemsetAsync(a.pair_ctl, 0, 3 * sizeof(int), s);
  {
    const int tiles = ((a.W + TW - 1) / TW) * ((a.H + TH - 1) / TH);
    bin_pairs_kernel<<<dim3(a.batch, max(1, min(16, (tiles + BIN_THREADS - 1) / BIN_THREADS))), BIN_THREADS, 0, s>>>(a);  // 64 tiles (and their share of the background rows) per block
  }
  return 1;
}

static int raster_grid() {  // resident blocks of the device for the persistent raster kernel (OFDG_RASTER_BLOCKS_PER_SM: fewer)
  static int raster_blocks = 0;
  if (!raster_blocks) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raster_pairs_kernel<false>, RASTER_THREADS, 0);
    if (const char* t = std::getenv("OFDG_RASTER_BLOCKS_PER_SM")) per_sm = min(per_sm, max(1, std::atoi(t)));
    raster_blocks = max(1, sms * max(1, per_sm));
  }
  return raster_blocks;
}

int launch_raster_pairs(const RenderArgs& a, cudaStream_t s) {
  if (a.n_fields > 0) raster_pairs_kernel<true><<<raster_grid(), RASTER_THREADS, 0, s>>>(a);
  else raster_pairs_kernel<false><<<raster_grid(), RASTER_THREADS, 0, s>>>(a);
  return 1;
}

int launch_render_split(const RenderArgs& a, cudaStream_t s, cudaEvent_t before_shade, int done) {
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH;
  dim3 grid(tiles_x * tiles_y, a.batch);
  const bool extra = a.flow_bw || a.top_id0 || a.top_id1 || a.ids8;
  int n = 0;
  if (done < 1) n += launch_bin_pairs(a, s);
  if (done < 2) n += launch_raster_pairs(a, s);
  if (before_shade) cudaEventRecord(before_shade, s);
  if (a.n_fields > 0) {
    if (extra) shade_kernel<true, true><<<grid, RENDER_THREADS, 0, s>>>(a);
    else shade_kernel<true, false><<<grid, RENDER_THREADS, 0, s>>>(a);
  } else {
    if (extra) shade_kernel<false, true><<<grid, RENDER_THREADS, 0, s>>>(a);
    else shade_kernel<false, false><<<grid, RENDER_THREADS, 0, s>>>(a);
  }
  ++n;
  if (!a.occlusion) return n;
  const size_t P = (size_t)a.W * a.H;
  occlusion_kernel<<<dim3((unsigned)((P + 255) / 256), a.batch), 256, 0, s>>>(a);
  return n + 1;
}

void launch_planar_to_rgbx(const uint8_t* planar, uchar4* out, int n, int w, int h, cudaStream_t s) {
  const size_t plane = (size_t)w * h;
  planar_to_rgbx_kernel<<<1184, 256, 0, s>>>(planar, out, plane * n, plane);
}
void launch_rgbx_to_planar(const uchar4* in, uint8_t* planar, int w, int h, cudaStream_t s) {
  rgbx_to_planar_kernel<<<592, 256, 0, s>>>(in, planar, (size_t)w * h);
}
void launch_synth_textures(uchar4* out, int n, int w, int h, uint64_t seed, int first_index, cudaStream_t s) {
  synth_textures_kernel<<<2368, 256, 0, s>>>(out, n, w, h, seed, first_index);
}
void launch_bg_to_planar(const uchar4* bg, uint8_t* planar, int batch, int w2, int h2, cudaStream_t s) {
  const size_t plane = (size_t)w2 * h2;
  for (int b = 0; b < batch; ++b) rgbx_to_planar_kernel<<<592, 256, 0, s>>>(bg + b * plane, planar + b * 3 * plane, plane);
}

}  // namespace ofdg
