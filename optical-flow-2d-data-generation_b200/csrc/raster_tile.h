// Tile-local restatement of AGG 2.4's scanline cell accumulation (rasterizer_cells_aa::line /
// render_hline) in closed form, so that any (edge, tile) pair can be evaluated
// independently of every other one:
//   * the x position where an edge leaves pixel row j is  xa + floor((p0 + 256*j*dx) / dy),
//   * the y position where a row segment leaves cell j is y1 + floor((p0 + 256*j*dy) / dx),
// which is what AGG's remainder-carrying DDAs compute step by step (floor division, positive
// divisor). Cells left of the tile only contribute their cover to the row's carry-in (covers
// telescope, no loop); cells right of the tile are irrelevant; clipping is off in the reference, so
// geometry left of x = 0 still feeds the carry.
//
// Compiled for the device (render.cu: atomics into shared memory) and for the host
// (api.cu: ofdg_debug_raster_host, exercised by the CPU tests against the oracle's sequential AGG
// port).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OFDG_HD __host__ __device__ __forceinline__
#define OFDG_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define OFDG_HD inline
#define OFDG_HD_NOINLINE inline
#endif

namespace ofdg {

#ifndef OFDG_TILE_ROWS
#define OFDG_TILE_ROWS 8
#endif
constexpr int TW = 128;             // tile width  = 32 lanes x 4 pixels
constexpr int TH = OFDG_TILE_ROWS;  // tile height = one warp per row

template <bool kDevice>
struct Acc;
template <>
struct Acc<false> {
  static OFDG_HD void add(int* p, int v) { *p += v; }
};
#if defined(__CUDACC__)
template <>
struct Acc<true> {
  static __device__ __forceinline__ void add(int* p, int v) { atomicAdd(p, v); }
};
#endif

OFDG_HD int rt_min(int a, int b) { return a < b ? a : b; }
OFDG_HD int rt_max(int a, int b) { return a > b ? a : b; }

// floor(a / b) and the matching non-negative remainder, b > 0. Out of line on the device (the emulated
// integer divide is ~40 instructions and there are five call sites); results come back in registers.
struct QuotRem {
  int q, r;
};
OFDG_HD_NOINLINE QuotRem floordivmod_qr(int a, int b) {
  QuotRem o;
  o.q = a / b;
  o.r = a - o.q * b;
  if (o.r < 0) { --o.q; o.r += b; }
  return o;
}
OFDG_HD void floordivmod(int a, int b, int& q, int& r) {
  const QuotRem o = floordivmod_qr(a, b);
  q = o.q;
  r = o.r;
}
// 64-bit numerator (|num| < 2^52), 32-bit positive divisor, quotient fits 32 bits. One correctly
// rounded double division plus an exact integer fix-up replaces the emulated 64-bit divide.
OFDG_HD_NOINLINE QuotRem floordivmod64_qr(long long num, int den) {
  long long qq = (long long)((double)num / (double)den);
  long long rr = num - qq * den;
  if (rr < 0) { --qq; rr += den; }
  if (rr < 0) { --qq; rr += den; }
  if (rr >= den) { ++qq; rr -= den; }
  QuotRem o;
  o.q = (int)qq;
  o.r = (int)rr;
  return o;
}
OFDG_HD void floordivmod64(long long num, int den, int& q, int& r) {
  const QuotRem o = floordivmod64_qr(num, den);
  q = o.q;
  r = o.r;
}

// render_hline(ey, x1, y1, x2, y2) scattered into one tile row. Out of line on the device: tile_edge
// calls it from two places and the render kernel has to stay inside the instruction cache.
template <bool kDevice>
OFDG_HD_NOINLINE void tile_hline(int* cover, int* area, int* carry, int tx0, int x1, int y1, int x2, int y2) {
  if (y1 == y2) return;
  const int ex1 = x1 >> 8, ex2 = x2 >> 8, fx1 = x1 & 255, fx2 = x2 & 255;
  const int dyv = y2 - y1;
  if (rt_max(ex1, ex2) < tx0) { Acc<kDevice>::add(carry, dyv); return; }  // entirely left: cover only
  if (rt_min(ex1, ex2) >= tx0 + TW) return;
  if (ex1 == ex2) {
    const int c = ex1 - tx0;
    Acc<kDevice>::add(cover + c, dyv);
    Acc<kDevice>::add(area + c, (fx1 + fx2) * dyv);
    return;
  }
  int dx = x2 - x1, p0, incr, first;
  if (dx > 0) { p0 = (256 - fx1) * dyv; incr = 1; first = 256; }
  else { p0 = fx1 * dyv; incr = -1; first = 0; dx = -dx; }
  const int ncell = (ex2 - ex1) * incr;  // cells j = 0..ncell along the walk, j-th cell = ex1 + incr*j
  // cumulative y after leaving cell j: C(j) = floor((p0 + 256*j*dyv) / dx) for j < ncell, C(ncell) = dyv
  int jlo, jhi;
  if (incr > 0) { jlo = rt_max(0, tx0 - ex1); jhi = rt_min(ncell, tx0 + TW - 1 - ex1); }
  else { jlo = rt_max(0, ex1 - (tx0 + TW - 1)); jhi = rt_min(ncell, ex1 - tx0); }
  int lift, rem;
  floordivmod(256 * dyv, dx, lift, rem);
  int prev = 0, c = 0, m = 0;  // C(jlo - 1) and the DDA state that produced it
  if (jlo > 0) {
    floordivmod(p0 + 256 * (jlo - 1) * dyv, dx, c, m);
    prev = c;
    if (incr > 0) Acc<kDevice>::add(carry, prev);  // the cells left of the tile (walk goes right)
  }
  for (int j = jlo; j <= jhi; ++j) {
    int cur;
    if (j >= ncell) cur = dyv;
    else if (j == 0) { floordivmod(p0, dx, c, m); cur = c; }
    else { c += lift; m += rem; if (m >= dx) { m -= dx; ++c; } cur = c; }
    const int d = cur - prev;
    prev = cur;
    int ar;
    if (j == 0) ar = (fx1 + first) * d;
    else if (j == ncell) ar = (fx2 + 256 - first) * d;
    else ar = 256 * d;
    const int cell = ex1 + incr * j - tx0;
    if (d | ar) { Acc<kDevice>::add(cover + cell, d); Acc<kDevice>::add(area + cell, ar); }
  }
  if (incr < 0 && ex2 < tx0) Acc<kDevice>::add(carry, dyv - prev);  // the cells left of the tile (walk goes left)
}

// Row range of edge (xa,ya)-(xb,yb) inside the tile (or its first `rows` rows): returns false when the edge cannot touch the tile's
// accumulators at all; `left` is set when it lies entirely left of the tile (carry-in only, no divisions).
OFDG_HD bool tile_edge_rows(int tx0, int ty0, int xa, int ya, int xb, int yb, int& rlo, int& rhi, bool& left, int rows = TH) {
  if (ya == yb) return false;
  const int ey1 = ya >> 8, ey2 = yb >> 8;
  rlo = rt_max(rt_min(ey1, ey2), ty0);
  rhi = rt_min(rt_max(ey1, ey2), ty0 + rows - 1);
  if (rlo > rhi) return false;
  if ((rt_min(xa, xb) >> 8) >= tx0 + TW) return false;
  left = (rt_max(xa, xb) >> 8) < tx0;
  return true;
}

// The part of rasterizer_cells_aa::line(xa, ya, xb, yb) that falls into pixel row r (ty0 <= r < ty0+TH),
// evaluated on its own: the x positions where the edge enters and leaves the row come from the closed form
//   X(j) = xa + floor((p0 + 256*j) * dx / dy),   j = index of the row along the edge,
// i.e. exactly the values AGG's remainder-carrying DDA reaches after j steps.
template <bool kDevice>
OFDG_HD void tile_edge_row(int* cover, int* area, int* carry, int tx0, int ty0, int r, int xa, int ya, int xb, int yb) {
  const int ey1 = ya >> 8, ey2 = yb >> 8, fy1 = ya & 255, fy2 = yb & 255;
  int* cv = cover + (r - ty0) * TW;
  int* ar = area + (r - ty0) * TW;
  int* cr = carry + (r - ty0);
  if (ey1 == ey2) {
    tile_hline<kDevice>(cv, ar, cr, tx0, xa, fy1, xb, fy2);
    return;
  }
  const bool down = yb > ya;
  const int ys = r == ey1 ? fy1 : (down ? 0 : 256), ye = r == ey2 ? fy2 : (down ? 256 : 0);
  if ((rt_max(xa, xb) >> 8) < tx0) {  // entirely left of the tile: only the y extent matters
    if (ye != ys) Acc<kDevice>::add(cr, ye - ys);
    return;
  }
  const int dx = xb - xa, dy = down ? yb - ya : ya - yb;
  const int p0 = down ? (256 - fy1) : fy1;
  const int j = down ? r - ey1 : ey1 - r;
  int xs = xa, xe = xb;
  int q = 0, m = 0;
  if (j > 0) {
    floordivmod64((long long)(p0 + 256LL * (j - 1)) * dx, dy, q, m);
    xs = xa + q;
  }
  if (r != ey2) {
    if (j == 0) floordivmod64((long long)p0 * dx, dy, q, m);
    else {
      int lift, rem;
      floordivmod(256 * dx, dy, lift, rem);
      q += lift; m += rem;
      if (m >= dy) { m -= dy; ++q; }
    }
    xe = xa + q;
  }
  tile_hline<kDevice>(cv, ar, cr, tx0, xs, ys, xe, ye);
}

// rasterizer_cells_aa::line(xa, ya, xb, yb) restricted to tile rows [ty0, ty0+TH) x columns [tx0, tx0+TW).
// cover/area: TH x TW ints, carry: TH ints.
template <bool kDevice>
OFDG_HD void tile_edge(int* cover, int* area, int* carry, int tx0, int ty0, int xa, int ya, int xb, int yb) {
  int rlo, rhi;
  bool left;
  if (!tile_edge_rows(tx0, ty0, xa, ya, xb, yb, rlo, rhi, left)) return;
  for (int r = rlo; r <= rhi; ++r) tile_edge_row<kDevice>(cover, area, carry, tx0, ty0, r, xa, ya, xb, yb);
}

// sweep_scanline + calculate_alpha for one pixel: `cum` = cover summed over all cells at or left of
// the pixel in its row, `area` = the pixel's own cell area. Arithmetic shift BEFORE abs (the winding
// sign changes the rounding), non-zero fill rule, clamp to 255.
OFDG_HD int coverage_alpha(int cum, int area) {
  int cv = (cum * 512 - area) >> 9;
  if (cv < 0) cv = -cv;
  return cv > 255 ? 255 : cv;
}
// pixfmt_gray8::blend_solid_hspan of colour 255 over a cleared buffer
OFDG_HD unsigned graylut(unsigned c) { return c == 255u ? 255u : (255u * ((255u * (c + 1u)) >> 8)) >> 8; }

}  // namespace ofdg
