// 2-D affine maps in IEEE double with exactly the operation order of AGG 2.4's
// agg::trans_affine (agg_trans_affine.h; the dependency is pinned by
// /root/reference/cmake/Dependencies.cmake:4-22 and is not vendored -- SURVEY App. B.2).
// The order matters: vertices go through iround(v * 256), so a 1-ulp difference can move a
// sub-pixel. Compile with -ffp-contract=off.
//
// Row-vector convention: (x, y) -> (x*sx + y*shx + tx, x*shy + y*sy + ty);
// `a.then(b)` applies a first, then b (AGG's a *= b).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define OFDG_AFFINE_FN __host__ __device__
#else
#define OFDG_AFFINE_FN
#endif

namespace ofdg {

struct Affine {
  double sx = 1, shy = 0, shx = 0, sy = 1, tx = 0, ty = 0;

  OFDG_AFFINE_FN static Affine rotation(double a) { return Affine{cos(a), sin(a), -sin(a), cos(a), 0.0, 0.0}; }
  OFDG_AFFINE_FN static Affine scaling(double s) { return Affine{s, 0.0, 0.0, s, 0.0, 0.0}; }
  OFDG_AFFINE_FN static Affine translation(double x, double y) { return Affine{1.0, 0.0, 0.0, 1.0, x, y}; }

  // trans_affine::multiply
  OFDG_AFFINE_FN Affine& then(const Affine& m) {
    double t0 = sx * m.sx + shy * m.shx;
    double t2 = shx * m.sx + sy * m.shx;
    double t4 = tx * m.sx + ty * m.shx + m.tx;
    shy = sx * m.shy + shy * m.sy;
    sy = shx * m.shy + sy * m.sy;
    ty = tx * m.shy + ty * m.sy + m.ty;
    sx = t0;
    shx = t2;
    tx = t4;
    return *this;
  }

  // trans_affine::invert
  OFDG_AFFINE_FN Affine inverse() const {
    Affine r = *this;
    double d = 1.0 / (r.sx * r.sy - r.shy * r.shx);
    double t0 = r.sy * d;
    r.sy = r.sx * d;
    r.shy = -r.shy * d;
    r.shx = -r.shx * d;
    double t4 = -r.tx * t0 - r.ty * r.shx;
    r.ty = -r.tx * r.shy - r.ty * r.sy;
    r.sx = t0;
    r.tx = t4;
    return r;
  }

  // trans_affine::transform
  OFDG_AFFINE_FN void apply(double* x, double* y) const {
    double tmp = *x;
    *x = tmp * sx + *y * shx + tx;
    *y = tmp * shy + *y * sy + ty;
  }

  OFDG_AFFINE_FN void store(double out[6]) const {
    out[0] = sx; out[1] = shy; out[2] = shx; out[3] = sy; out[4] = tx; out[5] = ty;
  }
};

// agg::iround
OFDG_AFFINE_FN static inline int iround(double v) { return int((v < 0.0) ? v - 0.5 : v + 0.5); }

}  // namespace ofdg
