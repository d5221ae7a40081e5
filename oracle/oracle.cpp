// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// A restatement, in the reference's own structure (whole-frame passes per object, sequential
// scanline rasteriser, CImg-style image ops), of the render path of
//   /root/reference/src/caffe/DataGenerator.cpp        ("DG.cpp" below)
//   /root/reference/src/caffe/WarpFields.cpp           (consumer side; producer in warpfields.cpp)
// plus the behaviour of the two un-vendored libraries it calls:
//   Anti-Grain Geometry 2.4 (pinned by /root/reference/cmake/Dependencies.cmake:4-22, MD5
//   863d9992fd83c5d40fe1c011501ecf0e) and CImg >= 2.0.0 (DataGenerator.h:50-51), restated from
//   their published algorithms as recorded in SURVEY.md App. B.
// PARITY UNPINNED: the reference has no golden vectors and cannot be built here; see oracle.h.
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile). No fast-math: float and
// double expressions are meant to round exactly like the reference's x86-64 SSE build.
#include "oracle.h"

#include "ofdg/augment.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace orc {

typedef unsigned char u8;
static thread_local std::string g_err;

// =================================================================================================
// CImg-like planar image: data[x + y*w + c*w*h]  
// =================================================================================================
template <class T>
struct Img {
  int w = 0, h = 0, c = 0;
  std::vector<T> d;
  const T* view = nullptr;  // non-owning (a pool texture); read-only
  Img() {}
  Img(int w_, int h_, int c_, T fill = T()) : w(w_), h(h_), c(c_), d((size_t)w_ * h_ * c_, fill) {}
  static Img<T> wrap(const T* p, int w_, int h_, int c_) { Img<T> r; r.w = w_; r.h = h_; r.c = c_; r.view = p; return r; }
  const T* data() const { return view ? view : d.data(); }
  T& at(int x, int y, int ch = 0) { return d[(size_t)x + (size_t)y * w + (size_t)ch * w * h]; }
  const T& at(int x, int y, int ch = 0) const { return data()[(size_t)x + (size_t)y * w + (size_t)ch * w * h]; }
};

static inline int cimg_mod_int(int x, int m) { return x >= 0 ? x % m : (x % m ? m + x % m : 0); }
static inline float cimg_mod_float(float x, float m) {
  const double dx = (double)x, dm = (double)m;
  return (float)(dx - dm * std::floor(dx / dm));
}

// get_crop(x0,y0,x1,y1, boundary 3 = mirror), inclusive corners
template <class T>
static Img<T> cimg_get_crop_mirror(const Img<T>& s, int x0, int y0, int x1, int y1) {
  const int nx0 = std::min(x0, x1), nx1 = std::max(x0, x1), ny0 = std::min(y0, y1), ny1 = std::max(y0, y1);
  Img<T> r(nx1 - nx0 + 1, ny1 - ny0 + 1, s.c);
  const int w2 = 2 * s.w, h2 = 2 * s.h;
  for (int ch = 0; ch < s.c; ++ch)
    for (int y = 0; y < r.h; ++y)
      for (int x = 0; x < r.w; ++x) {
        const int mx = cimg_mod_int(nx0 + x, w2), my = cimg_mod_int(ny0 + y, h2);
        r.at(x, y, ch) = s.at(mx < s.w ? mx : w2 - mx - 1, my < s.h ? my : h2 - my - 1, ch);
      }
  return r;
}

// get_shift(dx, dy, 0, 0, 3): mirror shift is a mirror crop at (-dx, -dy)
template <class T>
static Img<T> cimg_get_shift_mirror(const Img<T>& s, int dx, int dy) {
  return cimg_get_crop_mirror(s, -dx, -dy, s.w - dx - 1, s.h - dy - 1);
}

// _linear_atXY: Neumann (clamped) bilinear, float arithmetic
template <class T>
static inline float cimg_linear_neumann(const Img<T>& s, float fx, float fy, int ch) {
  const float nfx = fx <= 0 ? 0 : (fx >= s.w - 1 ? (float)(s.w - 1) : fx),
              nfy = fy <= 0 ? 0 : (fy >= s.h - 1 ? (float)(s.h - 1) : fy);
  const unsigned int x = (unsigned int)nfx, y = (unsigned int)nfy;
  const float dx = nfx - x, dy = nfy - y;
  const unsigned int nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
  const float Icc = (float)s.at(x, y, ch), Inc = (float)s.at(nx, y, ch), Icn = (float)s.at(x, ny, ch),
              Inn = (float)s.at(nx, ny, ch);
  return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
}

// linear_atXY(fx, fy, z, c, out_value): Dirichlet bilinear. NaN coordinates behave like the
// reference's x86 build: every tap is out of range and the NaN result converts to 0.
template <class T>
static inline float cimg_linear_dirichlet(const Img<T>& s, float fx, float fy, int ch, T out_value) {
  if (std::isnan(fx) || std::isnan(fy)) return std::nanf("");
  // float -> int conversions saturate here instead of being UB; far-out coordinates only
  // ever select out_value, so the result is the same as the reference's.
  auto toi = [](float v) { return v >= 2147483520.f ? 2147483520 : (v <= -2147483520.f ? -2147483520 : (int)v); };
  const int x = toi(fx) - (fx >= 0 ? 0 : 1), nx = x + 1, y = toi(fy) - (fy >= 0 ? 0 : 1), ny = y + 1;
  const float dx = fx - x, dy = fy - y;
  auto tap = [&](int px, int py) -> float {
    return (px < 0 || py < 0 || px >= s.w || py >= s.h) ? (float)out_value : (float)s.at(px, py, ch);
  };
  const float Icc = tap(x, y), Inc = tap(nx, y), Icn = tap(x, ny), Inn = tap(nx, ny);
  return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
}
static inline u8 to_u8_trunc(float v) { return std::isnan(v) ? (u8)0 : (u8)(int)v; }

// get_rotate(angle_deg, interpolation 1 = linear, boundary 3 = mirror)
static Img<u8> cimg_get_rotate_linear_mirror(const Img<u8>& s, float angle) {
  const float nangle = cimg_mod_float(angle, 360.0f);
  if (cimg_mod_float(nangle, 90.0f) == 0) {
    const int q = (int)nangle;
    if (q == 90 || q == 180 || q == 270) throw std::runtime_error("oracle: orthogonal rotations other than 0 are not restated");
    return s;  // +*this
  }
  const float rad = (float)(nangle * 3.14159265358979323846 / 180.0), ca = (float)std::cos(rad), sa = (float)std::sin(rad),
              ux = std::fabs((unsigned)(s.w - 1) * ca), uy = std::fabs((unsigned)(s.w - 1) * sa),
              vx = std::fabs((unsigned)(s.h - 1) * sa), vy = std::fabs((unsigned)(s.h - 1) * ca),
              w2 = 0.5f * (unsigned)(s.w - 1), h2 = 0.5f * (unsigned)(s.h - 1);
  Img<u8> r((int)std::floor((1 + ux + vx) + 0.5f), (int)std::floor((1 + uy + vy) + 0.5f), s.c);
  const float rw2 = 0.5f * (unsigned)(r.w - 1), rh2 = 0.5f * (unsigned)(r.h - 1);
  const float ww = 2.0f * s.w, hh = 2.0f * s.h;
  for (int ch = 0; ch < s.c; ++ch)
    for (int y = 0; y < r.h; ++y)
      for (int x = 0; x < r.w; ++x) {
        const float xc = x - rw2, yc = y - rh2, mx = cimg_mod_float(w2 + xc * ca + yc * sa, ww),
                    my = cimg_mod_float(h2 - xc * sa + yc * ca, hh);
        r.at(x, y, ch) = (u8)cimg_linear_neumann(s, mx < s.w ? mx : ww - mx - 1, my < s.h ? my : hh - my - 1, ch);
      }
  return r;
}

// get_resize(sx, sy, -100, -100, interpolation 3 = linear, boundary 0): separable; an axis that
// shrinks falls back to interpolation 2 (moving average); every pass stores T.
template <class T>
static Img<T> cimg_resize_axis(const Img<T>& s, int n, bool along_x) {
  const int len = along_x ? s.w : s.h;
  if (n == len) return s;
  Img<T> r(along_x ? n : s.w, along_x ? s.h : n, s.c);
  auto src = [&](int i, int j, int ch) -> T { return along_x ? s.at(i, j, ch) : s.at(j, i, ch); };
  auto dst = [&](int i, int j, int ch) -> T& { return along_x ? r.at(i, j, ch) : r.at(j, i, ch); };
  const int other = along_x ? s.h : s.w;
  if (len == 1) {  // nearest
    for (int ch = 0; ch < s.c; ++ch)
      for (int j = 0; j < other; ++j)
        for (int i = 0; i < n; ++i) dst(i, j, ch) = src(0, j, ch);
    return r;
  }
  if (len > n) {  // moving average
    std::vector<float> tmp((size_t)n * other * s.c, 0.f);
    auto T_ = [&](int t, int j, int ch) -> float& { return tmp[(size_t)t + (size_t)n * (j + (size_t)other * ch)]; };
    for (unsigned int a = (unsigned)len * n, b = len, c = n, sidx = 0, t = 0; a;) {
      const unsigned int dd = std::min(b, c);
      a -= dd; b -= dd; c -= dd;
      for (int ch = 0; ch < s.c; ++ch)
        for (int j = 0; j < other; ++j) T_(t, j, ch) += (float)src(sidx, j, ch) * dd;
      if (!b) {
        for (int ch = 0; ch < s.c; ++ch)
          for (int j = 0; j < other; ++j) T_(t, j, ch) /= (unsigned)len;
        ++t;
        b = len;
      }
      if (!c) { ++sidx; c = n; }
    }
    for (int ch = 0; ch < s.c; ++ch)
      for (int j = 0; j < other; ++j)
        for (int i = 0; i < n; ++i) dst(i, j, ch) = (T)T_(i, j, ch);
    return r;
  }
  // linear, growing
  const double f = n > 1 ? (len - 1.0) / (n - 1) : 0;
  std::vector<unsigned int> off(n);
  std::vector<double> foff(n);
  double curr = 0, old = 0;
  for (int i = 0; i < n; ++i) {
    foff[i] = curr - (unsigned int)curr;
    old = curr;
    curr = std::min(len - 1.0, curr + f);
    off[i] = (unsigned int)curr - (unsigned int)old;
  }
  for (int ch = 0; ch < s.c; ++ch)
    for (int j = 0; j < other; ++j) {
      int p = 0;
      for (int i = 0; i < n; ++i) {
        const double alpha = foff[i];
        const T val1 = src(p, j, ch), val2 = p < len - 1 ? src(p + 1, j, ch) : val1;
        dst(i, j, ch) = (T)((1 - alpha) * val1 + alpha * val2);
        p += off[i];
      }
    }
  return r;
}
template <class T>
static Img<T> cimg_get_resize_linear(const Img<T>& s, int sx, int sy) {
  return cimg_resize_axis(cimg_resize_axis(s, sx, true), sy, false);
}

// Texture::getRandomizedCrop, DG.cpp:87-109
static Img<u8> randomized_crop(const Img<u8>& tex, int tex_w, int tex_h, float angle, float zoom, int x_shift, int y_shift) {
  const int width = tex.w, height = tex.h;
  if (width >= tex_w && height >= tex_h) {
    Img<u8> r = cimg_get_rotate_linear_mirror(cimg_get_shift_mirror(tex, x_shift, y_shift), angle);
    r = cimg_get_crop_mirror(r, width / 2 - tex_w / 2, height / 2 - tex_h / 2, (int)(width / 2 - tex_w / 2 + tex_w / zoom - 1),
                             (int)(height / 2 - tex_h / 2 + tex_h / zoom - 1));
    return cimg_get_resize_linear(r, tex_w, tex_h);
  }
  Img<u8> r = cimg_get_rotate_linear_mirror(cimg_get_shift_mirror(tex, x_shift, y_shift), angle);
  return cimg_get_resize_linear(r, tex_w, tex_h);
}

// CImg::draw_image(x0=0, y0=0, sprite, mask, opacity=1, mask_max=255) for same-size images,
// the one-channel mask being cycled over the sprite's channels. DG.cpp:782,792
static void cimg_draw_image_masked(Img<u8>& dst, const Img<u8>& sprite, const Img<u8>& mask) {
  if (sprite.w != mask.w || sprite.h != mask.h) throw std::runtime_error("draw_image: sprite and mask differ");
  const float opacity = 1, mask_max_value = 255;
  const int n = dst.w * dst.h;
  for (int ch = 0; ch < dst.c; ++ch) {
    u8* ptrd = &dst.d[(size_t)ch * n];
    const u8* ptrs = &sprite.d[(size_t)ch * n];
    const u8* ptrm = mask.d.data();
    for (int i = 0; i < n; ++i) {
      const float mopacity = (float)(ptrm[i] * opacity), nopacity = std::fabs(mopacity),
                  copacity = mask_max_value - std::max(mopacity, 0.f);
      ptrd[i] = (u8)((nopacity * ptrs[i] + ptrd[i] * copacity) / mask_max_value);
    }
  }
}

// =================================================================================================
// AGG 2.4 restated
// =================================================================================================
static const double kPi = 3.14159265358979323846;
static inline int iround(double v) { return int((v < 0.0) ? v - 0.5 : v + 0.5); }
static inline unsigned uround(double v) { return unsigned(v + 0.5); }

struct TransAffine {  // agg::trans_affine
  double sx = 1, shy = 0, shx = 0, sy = 1, tx = 0, ty = 0;
  TransAffine() {}
  TransAffine(double a, double b, double c, double d, double e, double f) : sx(a), shy(b), shx(c), sy(d), tx(e), ty(f) {}
  static TransAffine rotation(double a) { return TransAffine(std::cos(a), std::sin(a), -std::sin(a), std::cos(a), 0.0, 0.0); }
  static TransAffine scaling(double s) { return TransAffine(s, 0.0, 0.0, s, 0.0, 0.0); }
  static TransAffine translation(double x, double y) { return TransAffine(1.0, 0.0, 0.0, 1.0, x, y); }
  const TransAffine& multiply(const TransAffine& m) {
    double t0 = sx * m.sx + shy * m.shx;
    double t2 = shx * m.sx + sy * m.shx;
    double t4 = tx * m.sx + ty * m.shx + m.tx;
    shy = sx * m.shy + shy * m.sy;
    sy = shx * m.shy + sy * m.sy;
    ty = tx * m.shy + ty * m.sy + m.ty;
    sx = t0; shx = t2; tx = t4;
    return *this;
  }
  const TransAffine& operator*=(const TransAffine& m) { return multiply(m); }
  TransAffine operator*(const TransAffine& m) const { return TransAffine(*this).multiply(m); }
  const TransAffine& invert() {
    double d = 1.0 / (sx * sy - shy * shx);
    double t0 = sy * d;
    sy = sx * d;
    shy = -shy * d;
    shx = -shx * d;
    double t4 = -tx * t0 - ty * shx;
    ty = -tx * shy - ty * sy;
    sx = t0; tx = t4;
    return *this;
  }
  void transform(double* x, double* y) const {
    double tmp = *x;
    *x = tmp * sx + *y * shx + tx;
    *y = tmp * shy + *y * sy + ty;
  }
};

enum { cmd_stop = 0, cmd_move_to = 1, cmd_line_to = 2, cmd_curve3 = 3, cmd_end_poly_close = 0x4F };
static inline bool is_vertex(unsigned c) { return c >= cmd_move_to && c < 0x0F; }

struct EllipseVS {  // agg::ellipse
  double x = 0, y = 0, rx = 1, ry = 1;
  unsigned num = 4, step = 0;
  void init(double x_, double y_, double rx_, double ry_, unsigned n) { x = x_; y = y_; rx = rx_; ry = ry_; num = n; step = 0; }
  void rewind() { step = 0; }
  unsigned vertex(double* px, double* py) {
    if (step == num) { ++step; return cmd_end_poly_close; }
    if (step > num) return cmd_stop;
    double angle = double(step) / double(num) * 2.0 * kPi;
    *px = x + std::cos(angle) * rx;
    *py = y + std::sin(angle) * ry;
    step++;
    return ((step == 1) ? cmd_move_to : cmd_line_to);
  }
};

struct PathStorage {  // agg::path_storage (vertex list part)
  struct V { double x, y; unsigned cmd; };
  std::vector<V> v;
  size_t it = 0;
  void remove_all() { v.clear(); it = 0; }
  void move_to(double x, double y) { v.push_back({x, y, cmd_move_to}); }
  void line_to(double x, double y) { v.push_back({x, y, cmd_line_to}); }
  void curve3(double xc, double yc, double xt, double yt) { v.push_back({xc, yc, cmd_curve3}); v.push_back({xt, yt, cmd_curve3}); }
  void close_polygon() { if (!v.empty() && is_vertex(v.back().cmd)) v.push_back({0.0, 0.0, cmd_end_poly_close}); }
  void rewind() { it = 0; }
  unsigned vertex(double* x, double* y) {
    if (it >= v.size()) return cmd_stop;
    *x = v[it].x; *y = v[it].y;
    return v[it++].cmd;
  }
};

template <class VS>
struct ConvTransform {  // agg::conv_transform
  VS* src; const TransAffine* tr;
  ConvTransform(VS& s, const TransAffine& t) : src(&s), tr(&t) {}
  void rewind() { src->rewind(); }
  unsigned vertex(double* x, double* y) {
    unsigned cmd = src->vertex(x, y);
    if (is_vertex(cmd)) tr->transform(x, y);
    return cmd;
  }
};

struct Curve3Div {  // agg::curve3_div, approximation_scale 1, angle_tolerance 0
  struct P { double x, y; };
  std::vector<P> pts;
  size_t count = 0;
  double dist_tol_sq = 0.25;
  void reset() { pts.clear(); count = 0; }
  void init(double x1, double y1, double x2, double y2, double x3, double y3) {
    pts.clear();
    dist_tol_sq = 0.5 / 1.0;
    dist_tol_sq *= dist_tol_sq;
    pts.push_back({x1, y1});
    recursive_bezier(x1, y1, x2, y2, x3, y3, 0);
    pts.push_back({x3, y3});
    count = 0;
  }
  static double sqd(double x1, double y1, double x2, double y2) { double dx = x2 - x1, dy = y2 - y1; return dx * dx + dy * dy; }
  void recursive_bezier(double x1, double y1, double x2, double y2, double x3, double y3, unsigned level) {
    if (level > 32) return;
    double x12 = (x1 + x2) / 2, y12 = (y1 + y2) / 2, x23 = (x2 + x3) / 2, y23 = (y2 + y3) / 2;
    double x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2;
    double dx = x3 - x1, dy = y3 - y1;
    double d = std::fabs(((x2 - x3) * dy - (y2 - y3) * dx));
    double da;
    if (d > 1e-30) {
      if (d * d <= dist_tol_sq * (dx * dx + dy * dy)) {
        pts.push_back({x123, y123});  // angle_tolerance < epsilon
        return;
      }
    } else {
      da = dx * dx + dy * dy;
      if (da == 0) {
        d = sqd(x1, y1, x2, y2);
      } else {
        d = ((x2 - x1) * dx + (y2 - y1) * dy) / da;
        if (d > 0 && d < 1) return;
        if (d <= 0) d = sqd(x2, y2, x1, y1);
        else if (d >= 1) d = sqd(x2, y2, x3, y3);
        else d = sqd(x2, y2, x1 + d * dx, y1 + d * dy);
      }
      if (d < dist_tol_sq) {
        pts.push_back({x2, y2});
        return;
      }
    }
    recursive_bezier(x1, y1, x12, y12, x123, y123, level + 1);
    recursive_bezier(x123, y123, x23, y23, x3, y3, level + 1);
  }
  unsigned vertex(double* x, double* y) {
    if (count >= pts.size()) return cmd_stop;
    const P& p = pts[count++];
    *x = p.x; *y = p.y;
    return (count == 1) ? cmd_move_to : cmd_line_to;
  }
};

template <class VS>
struct ConvCurve {  // agg::conv_curve (curve3 only; the reference never emits curve4)
  VS* src;
  double last_x = 0, last_y = 0;
  Curve3Div c3;
  explicit ConvCurve(VS& s) : src(&s) {}
  void rewind() { src->rewind(); last_x = last_y = 0; c3.reset(); }
  unsigned vertex(double* x, double* y) {
    if (c3.vertex(x, y) != cmd_stop) { last_x = *x; last_y = *y; return cmd_line_to; }
    double end_x = 0, end_y = 0;
    unsigned cmd = src->vertex(x, y);
    if (cmd == cmd_curve3) {
      src->vertex(&end_x, &end_y);
      c3.init(last_x, last_y, *x, *y, end_x, end_y);
      c3.vertex(x, y);  // move_to
      c3.vertex(x, y);  // first vertex of the curve
      cmd = cmd_line_to;
    }
    last_x = *x; last_y = *y;
    return cmd;
  }
};

// Pre-rounded 24.8 vertices (oracle_raster_fixed)
struct FixedVS {
  const int32_t* xy; int n; int i = 0;
  void rewind() { i = 0; }
};

struct Cell { int x, y, cover, area; };

// agg::scanline_u8 content for one y
struct Span { int x, len; size_t cover_off; };
struct Scanline {
  int y = 0, min_x = 0, last_x = 0x7FFFFFF0;
  std::vector<u8> covers;
  std::vector<Span> spans;
  void reset(int mn, int mx) { min_x = mn; covers.assign((size_t)(mx - mn + 3), 0); last_x = 0x7FFFFFF0; spans.clear(); }
  void reset_spans() { last_x = 0x7FFFFFF0; spans.clear(); }
  void add_cell(int x, unsigned cover) {
    x -= min_x;
    covers[x] = (u8)cover;
    if (x == last_x + 1) spans.back().len++;
    else spans.push_back({(int)(int16_t)(x + min_x), 1, (size_t)x});
    last_x = x;
  }
  void add_span(int x, unsigned len, unsigned cover) {
    x -= min_x;
    std::memset(&covers[x], (int)cover, len);
    if (x == last_x + 1) spans.back().len += (int)len;
    else spans.push_back({(int)(int16_t)(x + min_x), (int)len, (size_t)x});
    last_x = x + (int)len - 1;
  }
};

// agg::rasterizer_scanline_aa<rasterizer_sl_clip_int> over rasterizer_cells_aa<cell_aa>;
// clipping disabled (clip_box is never called by the reference), non-zero fill, auto_close.
struct Rasterizer {
  enum { shift = 8, scale = 256, mask = 255 };
  std::vector<Cell> cells;
  Cell curr{0x7FFFFFFF, 0x7FFFFFFF, 0, 0};
  int min_x = 0x7FFFFFFF, min_y = 0x7FFFFFFF, max_x = -0x7FFFFFFF, max_y = -0x7FFFFFFF;
  int start_x = 0, start_y = 0, x1 = 0, y1 = 0;
  enum { st_initial, st_move_to, st_line_to, st_closed } status = st_initial;
  bool sorted = false;
  unsigned gamma[256];
  size_t sweep_i = 0;
  int scan_y = 0;

  Rasterizer() { for (int i = 0; i < 256; ++i) gamma[i] = i; }
  void reset() {
    cells.clear();
    curr = Cell{0x7FFFFFFF, 0x7FFFFFFF, 0, 0};
    min_x = min_y = 0x7FFFFFFF; max_x = max_y = -0x7FFFFFFF;
    status = st_initial; sorted = false;
  }
  template <class F> void set_gamma(F f) { for (int i = 0; i < 256; ++i) gamma[i] = uround(f(double(i) / 255) * 255); }

  void add_curr_cell() { if (curr.area | curr.cover) cells.push_back(curr); }
  void set_curr_cell(int x, int y) {
    if (curr.x != x || curr.y != y) { add_curr_cell(); curr.x = x; curr.y = y; curr.cover = 0; curr.area = 0; }
  }
  void render_hline(int ey, int xa, int ya, int xb, int yb) {
    int ex1 = xa >> shift, ex2 = xb >> shift, fx1 = xa & mask, fx2 = xb & mask;
    int delta, p, first, dx, incr, lift, mod, rem;
    if (ya == yb) { set_curr_cell(ex2, ey); return; }
    if (ex1 == ex2) { delta = yb - ya; curr.cover += delta; curr.area += (fx1 + fx2) * delta; return; }
    p = (scale - fx1) * (yb - ya); first = scale; incr = 1; dx = xb - xa;
    if (dx < 0) { p = fx1 * (yb - ya); first = 0; incr = -1; dx = -dx; }
    delta = p / dx; mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    curr.cover += delta; curr.area += (fx1 + first) * delta;
    ex1 += incr; set_curr_cell(ex1, ey); ya += delta;
    if (ex1 != ex2) {
      p = scale * (yb - ya + delta); lift = p / dx; rem = p % dx;
      if (rem < 0) { lift--; rem += dx; }
      mod -= dx;
      while (ex1 != ex2) {
        delta = lift; mod += rem;
        if (mod >= 0) { mod -= dx; delta++; }
        curr.cover += delta; curr.area += scale * delta;
        ya += delta; ex1 += incr; set_curr_cell(ex1, ey);
      }
    }
    delta = yb - ya;
    curr.cover += delta; curr.area += (fx2 + scale - first) * delta;
  }
  void line(int xa, int ya, int xb, int yb) {
    enum { dx_limit = 16384 << shift };
    int dx = xb - xa;
    if (dx >= dx_limit || dx <= -dx_limit) {
      int cx = (xa + xb) >> 1, cy = (ya + yb) >> 1;
      line(xa, ya, cx, cy); line(cx, cy, xb, yb);
      return;
    }
    int dy = yb - ya;
    int ex1 = xa >> shift, ex2 = xb >> shift, ey1 = ya >> shift, ey2 = yb >> shift, fy1 = ya & mask, fy2 = yb & mask;
    int x_from, x_to, p, rem, mod, lift, delta, first, incr;
    if (ex1 < min_x) min_x = ex1; if (ex1 > max_x) max_x = ex1;
    if (ey1 < min_y) min_y = ey1; if (ey1 > max_y) max_y = ey1;
    if (ex2 < min_x) min_x = ex2; if (ex2 > max_x) max_x = ex2;
    if (ey2 < min_y) min_y = ey2; if (ey2 > max_y) max_y = ey2;
    set_curr_cell(ex1, ey1);
    if (ey1 == ey2) { render_hline(ey1, xa, fy1, xb, fy2); return; }
    incr = 1;
    if (dx == 0) {
      int ex = xa >> shift, two_fx = (xa - (ex << shift)) << 1, area;
      first = scale;
      if (dy < 0) { first = 0; incr = -1; }
      x_from = xa;
      delta = first - fy1;
      curr.cover += delta; curr.area += two_fx * delta;
      ey1 += incr; set_curr_cell(ex, ey1);
      delta = first + first - scale; area = two_fx * delta;
      while (ey1 != ey2) { curr.cover = delta; curr.area = area; ey1 += incr; set_curr_cell(ex, ey1); }
      delta = fy2 - scale + first;
      curr.cover += delta; curr.area += two_fx * delta;
      return;
    }
    p = (scale - fy1) * dx; first = scale;
    if (dy < 0) { p = fy1 * dx; first = 0; incr = -1; dy = -dy; }
    delta = p / dy; mod = p % dy;
    if (mod < 0) { delta--; mod += dy; }
    x_from = xa + delta;
    render_hline(ey1, xa, fy1, x_from, first);
    ey1 += incr; set_curr_cell(x_from >> shift, ey1);
    if (ey1 != ey2) {
      p = scale * dx; lift = p / dy; rem = p % dy;
      if (rem < 0) { lift--; rem += dy; }
      mod -= dy;
      while (ey1 != ey2) {
        delta = lift; mod += rem;
        if (mod >= 0) { mod -= dy; delta++; }
        x_to = x_from + delta;
        render_hline(ey1, x_from, scale - first, x_to, first);
        x_from = x_to;
        ey1 += incr; set_curr_cell(x_from >> shift, ey1);
      }
    }
    render_hline(ey1, x_from, scale - first, xb, fy2);
  }
  // rasterizer_scanline_aa vertex feed
  void close_polygon() { if (status == st_line_to) { line(x1, y1, start_x, start_y); x1 = start_x; y1 = start_y; status = st_closed; } }
  void move_to_fixed(int x, int y) {
    if (sorted) reset();
    close_polygon();
    start_x = x1 = x; start_y = y1 = y; status = st_move_to;
  }
  void line_to_fixed(int x, int y) { line(x1, y1, x, y); x1 = x; y1 = y; status = st_line_to; }
  void move_to_d(double x, double y) { move_to_fixed(iround(x * scale), iround(y * scale)); }
  void line_to_d(double x, double y) { line_to_fixed(iround(x * scale), iround(y * scale)); }
  template <class VS> void add_path(VS& vs) {
    double x = 0, y = 0; unsigned cmd;
    vs.rewind();
    if (sorted) reset();
    while ((cmd = vs.vertex(&x, &y)) != cmd_stop) {
      if (cmd == cmd_move_to) move_to_d(x, y);
      else if (is_vertex(cmd)) line_to_d(x, y);
      else if (cmd == cmd_end_poly_close) close_polygon();
    }
  }
  bool rewind_scanlines() {
    close_polygon();
    if (!sorted) {
      add_curr_cell();
      curr = Cell{0x7FFFFFFF, 0x7FFFFFFF, 0, 0};
      std::stable_sort(cells.begin(), cells.end(), [](const Cell& a, const Cell& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
      sorted = true;
    }
    if (cells.empty()) return false;
    sweep_i = 0; scan_y = min_y;
    return true;
  }
  unsigned calculate_alpha(int area) const {
    int cover = area >> (shift * 2 + 1 - 8);
    if (cover < 0) cover = -cover;
    if (cover > 255) cover = 255;
    return gamma[cover];
  }
  bool sweep_scanline(Scanline& sl) {
    for (;;) {
      if (scan_y > max_y) return false;
      sl.reset_spans();
      size_t j = sweep_i;
      while (j < cells.size() && cells[j].y == scan_y) ++j;
      size_t num_cells = j - sweep_i;
      const Cell* cp = cells.data() + sweep_i;
      sweep_i = j;
      int cover = 0;
      while (num_cells) {
        const Cell* cur = cp;
        int x = cur->x, area = cur->area;
        unsigned alpha;
        cover += cur->cover;
        while (--num_cells) {
          cur = ++cp;
          if (cur->x != x) break;
          area += cur->area; cover += cur->cover;
        }
        if (area) {
          alpha = calculate_alpha((cover << (shift + 1)) - area);
          if (alpha) sl.add_cell(x, alpha);
          x++;
        }
        if (num_cells && cur->x > x) {
          alpha = calculate_alpha(cover << (shift + 1));
          if (alpha) sl.add_span(x, cur->x - x, alpha);
        }
      }
      if (!sl.spans.empty()) break;
      ++scan_y;
    }
    sl.y = scan_y;
    ++scan_y;
    return true;
  }
};

// renderer_scanline_aa_solid<renderer_base<pixfmt_gray8>> with color gray8(255)
static void render_scanlines_gray8(Rasterizer& ras, u8* buf, int W, int H) {
  if (!ras.rewind_scanlines()) return;
  Scanline sl;
  sl.reset(ras.min_x, ras.max_x);
  while (ras.sweep_scanline(sl)) {
    const int y = sl.y;
    if (y < 0 || y > H - 1) continue;  // renderer_base clip box
    for (const Span& sp : sl.spans) {
      int x = sp.x, len = sp.len;
      const u8* covers = &sl.covers[sp.cover_off];
      if (x > W - 1) continue;
      if (x < 0) { len += x; if (len <= 0) continue; covers -= x; x = 0; }
      if (x + len > W) { len = W - x; if (len <= 0) continue; }
      u8* p = buf + (size_t)y * W + x;
      do {  // pixfmt_gray8::blend_solid_hspan, c.v = c.a = 255
        unsigned alpha = (255u * (unsigned(*covers) + 1)) >> 8;
        if (alpha == 255) *p = 255;
        else *p = (u8)((((255 - int(*p)) * int(alpha)) + (int(*p) << 8)) >> 8);
        ++p; ++covers;
      } while (--len);
    }
  }
}

// dda2_line_interpolator
struct Dda2 {
  int cnt, lft, rem, mod, y;
  Dda2() : cnt(1), lft(0), rem(0), mod(0), y(0) {}
  Dda2(int y1, int y2, int count) : cnt(count <= 0 ? 1 : count), lft((y2 - y1) / cnt), rem((y2 - y1) % cnt), mod(rem), y(y1) {
    if (mod <= 0) { mod += count; rem += count; lft--; }
    mod -= count;
  }
  void operator++() { mod += rem; y += lft; if (mod > 0) { mod -= cnt; y++; } }
};
struct WrapReflect {  // agg::wrap_mode_reflect
  unsigned size, size2, add, value;
  explicit WrapReflect(unsigned s) : size(s), size2(s * 2), add(size2 * (0x3FFFFFFF / size2)), value(0) {}
  unsigned operator()(int v) { value = (unsigned(v) + add) % size2; if (value >= size) return size2 - value - 1; return value; }
  unsigned operator++() { ++value; if (value >= size2) value = 0; if (value >= size) return size2 - value - 1; return value; }
};

// getTransformedTexture, DG.cpp:168-231
static Img<u8> get_transformed_texture(const Img<u8>& input, const TransAffine& tf_ref) {
  const int tex_W = input.w, tex_H = input.h;
  // permute_axes("CXYZ"): planar -> RGBRGB...
  std::vector<u8> src((size_t)tex_W * tex_H * 3), out((size_t)tex_W * tex_H * 3, 0);
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < tex_H; ++y)
      for (int x = 0; x < tex_W; ++x) src[((size_t)y * tex_W + x) * 3 + c] = input.at(x, y, c);
  TransAffine image_mtx = tf_ref;
  image_mtx.invert();
  Rasterizer ras;
  PathStorage path;
  path.move_to(0, 0); path.line_to(tex_W, 0); path.line_to(tex_W, tex_H); path.line_to(0, tex_H); path.close_polygon();
  ras.add_path(path);
  if (ras.rewind_scanlines()) {
    Scanline sl;
    sl.reset(ras.min_x, ras.max_x);
    WrapReflect wrap_x(tex_W), wrap_y(tex_H);
    std::vector<u8> span_rgb;
    while (ras.sweep_scanline(sl)) {
      const int y = sl.y;
      for (const Span& sp : sl.spans) {  // render_scanline_aa
        int x = sp.x, len = sp.len;
        const u8* covers = &sl.covers[sp.cover_off];
        span_rgb.assign((size_t)len * 3, 0);
        {  // span_image_filter_rgb_bilinear::generate(span, x, y, len)
          double tx = x + 0.5, ty = y + 0.5;  // span_interpolator_linear::begin
          image_mtx.transform(&tx, &ty);
          int ix1 = iround(tx * 256), iy1 = iround(ty * 256);
          tx = (x + 0.5) + len; ty = y + 0.5;
          image_mtx.transform(&tx, &ty);
          int ix2 = iround(tx * 256), iy2 = iround(ty * 256);
          Dda2 li_x(ix1, ix2, len), li_y(iy1, iy2, len);
          for (int i = 0; i < len; ++i) {
            int x_hr = li_x.y - 128, y_hr = li_y.y - 128;
            int x_lr = x_hr >> 8, y_lr = y_hr >> 8;
            unsigned fg[3] = {256 * 256 / 2, 256 * 256 / 2, 256 * 256 / 2}, weight;
            x_hr &= 255; y_hr &= 255;
            const u8* row = &src[(size_t)wrap_y(y_lr) * tex_W * 3];  // span(x_lr, y_lr, 2)
            const u8* p = row + wrap_x(x_lr) * 3;
            weight = (256 - x_hr) * (256 - y_hr);
            fg[0] += weight * p[0]; fg[1] += weight * p[1]; fg[2] += weight * p[2];
            p = row + (++wrap_x) * 3;  // next_x
            weight = x_hr * (256 - y_hr);
            fg[0] += weight * p[0]; fg[1] += weight * p[1]; fg[2] += weight * p[2];
            row = &src[(size_t)(++wrap_y) * tex_W * 3];  // next_y
            p = row + wrap_x(x_lr) * 3;
            weight = (256 - x_hr) * y_hr;
            fg[0] += weight * p[0]; fg[1] += weight * p[1]; fg[2] += weight * p[2];
            p = row + (++wrap_x) * 3;  // next_x
            weight = x_hr * y_hr;
            fg[0] += weight * p[0]; fg[1] += weight * p[1]; fg[2] += weight * p[2];
            span_rgb[i * 3 + 0] = (u8)(fg[0] >> 16); span_rgb[i * 3 + 1] = (u8)(fg[1] >> 16); span_rgb[i * 3 + 2] = (u8)(fg[2] >> 16);
            ++li_x; ++li_y;
          }
        }
        // renderer_base::blend_color_hspan: clip, then pixfmt_rgb24 copy_or_blend with alpha 255
        if (y < 0 || y > tex_H - 1) continue;
        const u8* colors = span_rgb.data();
        if (x < 0) { int d = -x; len -= d; if (len <= 0) continue; covers += d; colors += d * 3; x = 0; }
        if (x + len > tex_W) { len = tex_W - x; if (len <= 0) continue; }
        u8* p = &out[((size_t)y * tex_W + x) * 3];
        for (int i = 0; i < len; ++i, p += 3, colors += 3) {
          unsigned alpha = (255u * (unsigned(covers[i]) + 1)) >> 8;
          if (alpha == 255) { p[0] = colors[0]; p[1] = colors[1]; p[2] = colors[2]; }
          else for (int c = 0; c < 3; ++c) p[c] = (u8)((((int(colors[c]) - int(p[c])) * int(alpha)) + (int(p[c]) << 8)) >> 8);
        }
      }
    }
  }
  // permute_axes("YZCX"): back to planar
  Img<u8> res(tex_W, tex_H, 3);
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < tex_H; ++y)
      for (int x = 0; x < tex_W; ++x) res.at(x, y, c) = out[((size_t)y * tex_W + x) * 3 + c];
  return res;
}

// applyWarpFieldToTexture, DG.cpp:237-252. iflow is 2-channel; the reference indexes the channel
// through the z slot, which lands on the channel plane because depth is 1.
static Img<u8> apply_warp_field(const Img<u8>& input, const Img<float>& iflow) {
  Img<u8> result(input.w, input.h, input.c);
  for (int c = 0; c < input.c; ++c)
    for (int y = 0; y < input.h; ++y)
      for (int x = 0; x < input.w; ++x)
        result.at(x, y, c) = to_u8_trunc(cimg_linear_dirichlet(input, x + iflow.at(x, y, 0), y + iflow.at(x, y, 1), c, (u8)0));
  return result;
}

// =================================================================================================
// Scene objects (DG.cpp:256-718)
// =================================================================================================
struct Ctx {
  int W, H, mode;
  bool use_aa, faithful;
  const oracle_config* cfg;
  const uint8_t* textures;
  const float* fields;
};

struct MovingObject {
  size_t ID = 0;
  int kind = 0;  // OFDG_OBJ_*; background is a polygon with is_bg
  bool is_bg = false, is_component = false;
  bool has_fields = false;
  Img<float> field, field_inv;
  std::vector<u8> mask_noAA[2], mask_AA[2], scratch;
  std::vector<Img<u8>> textures;
  TransAffine intrinsic, intrinsic_inv, motion, motion_inv;
  EllipseVS ellipse;
  PathStorage path;
  std::vector<std::unique_ptr<MovingObject>> components;
  std::vector<bool> component_modes;
  Rasterizer ras;
  const Ctx* ctx = nullptr;

  MovingObject(const Ctx* c, size_t id) : ID(id), ctx(c) {
    const size_t n = (size_t)c->W * c->H;
    scratch.assign(n, 0);
    for (int f = 0; f < 2; ++f) { mask_noAA[f].assign(n, 0); mask_AA[f].assign(n, 0); }
  }
  void set_intrinsic(float alpha, float xs, float ys) {  // DG.cpp:302-310
    intrinsic = TransAffine();
    intrinsic *= TransAffine::rotation(alpha);
    intrinsic *= TransAffine::translation(xs, ys);
    intrinsic_inv = intrinsic;
    intrinsic_inv.invert();
  }
  void set_motion(float alpha, float scale, float xs, float ys) {  // DG.cpp:312-322
    motion = TransAffine();
    motion *= TransAffine::rotation(alpha);
    motion *= TransAffine::scaling(scale);
    motion *= TransAffine::translation(xs, ys);
    motion_inv = motion;
    motion_inv.invert();
  }
  void add_background_motion(const TransAffine& bg_motion) {  // DG.cpp:324-335
    const int W = ctx->W, H = ctx->H;
    TransAffine bg_n = TransAffine::translation(-W / 2., -H / 2.);
    bg_n *= bg_motion;
    bg_n *= TransAffine::translation(W / 2., H / 2.);
    motion *= bg_n;
    motion_inv = motion;
    motion_inv.invert();
  }
  template <class VS> void draw(VS& vs, bool AA, unsigned frame_idx) {  // DG.cpp:351-368
    std::fill(scratch.begin(), scratch.end(), (u8)0);
    ras.reset();
    ras.add_path(vs);
    if (AA) ras.set_gamma([](double x) { return x; });
    else ras.set_gamma([](double x) { return (x < 0.5) ? 0.0 : 1.0; });
    render_scanlines_gray8(ras, scratch.data(), ctx->W, ctx->H);
    if (AA) mask_AA[frame_idx] = scratch; else mask_noAA[frame_idx] = scratch;
  }
  void warp_masks() {  // MovingObjectBase::renderMasks, DG.cpp:370-386
    if (!has_fields) return;
    for (int k = 0; k < 2; ++k) {
      std::vector<u8>& m = k == 0 ? mask_noAA[1] : mask_AA[1];
      Img<u8> tmp(ctx->W, ctx->H, 1);
      tmp.d = m;
      m = apply_warp_field(tmp, field_inv).d;
    }
  }
  void render_transformed_texture() {
    if (is_component) return;  // DG.cpp:546-560
    if (is_bg) {               // DG.cpp:665-682
      const int W = ctx->W, H = ctx->H;
      textures.push_back(get_transformed_texture(textures[0], TransAffine()));
      TransAffine tf = intrinsic_inv * motion * intrinsic;
      if (has_fields) textures.push_back(apply_warp_field(get_transformed_texture(textures[0], tf), field_inv));
      else textures.push_back(get_transformed_texture(textures[0], tf));
      textures[1] = cimg_get_crop_mirror(textures[1], (int)(W / 2.), (int)(H / 2.), (int)(W * 3. / 2. - 1), (int)(H * 3. / 2. - 1));
      textures[2] = cimg_get_crop_mirror(textures[2], (int)(W / 2.), (int)(H / 2.), (int)(W * 3. / 2. - 1), (int)(H * 3. / 2. - 1));
      return;
    }
    textures.push_back(get_transformed_texture(textures[0], TransAffine()));  // DG.cpp:337-349
    if (has_fields) textures.push_back(apply_warp_field(get_transformed_texture(textures[0], motion), field_inv));
    else textures.push_back(get_transformed_texture(textures[0], motion));
  }
  void render_masks() {
    const size_t n = (size_t)ctx->W * ctx->H;
    if (is_bg) {  // DG.cpp:684-690
      for (int f = 0; f < 2; ++f) { std::fill(mask_AA[f].begin(), mask_AA[f].end(), (u8)255); std::fill(mask_noAA[f].begin(), mask_noAA[f].end(), (u8)255); }
      return;
    }
    TransAffine save = intrinsic;
    save *= motion;
    if (kind == OFDG_OBJ_ELLIPSE) {  // DG.cpp:465-479
      ConvTransform<EllipseVS> e0(ellipse, intrinsic);
      draw(e0, true, 0); draw(e0, false, 0);
      ConvTransform<EllipseVS> e1(ellipse, save);
      draw(e1, true, 1); draw(e1, false, 1);
      warp_masks();
    } else if (kind == OFDG_OBJ_POLYGON) {  // DG.cpp:520-534
      ConvTransform<PathStorage> p0(path, intrinsic);
      ConvCurve<ConvTransform<PathStorage>> c0(p0);
      ConvTransform<PathStorage> p1(path, save);
      ConvCurve<ConvTransform<PathStorage>> c1(p1);
      draw(c0, true, 0); draw(c0, false, 0);
      draw(c1, true, 1); draw(c1, false, 1);
      warp_masks();
    } else {  // composite, DG.cpp:591-646
      for (int f = 0; f < 2; ++f) { std::fill(mask_AA[f].begin(), mask_AA[f].end(), (u8)0); std::fill(mask_noAA[f].begin(), mask_noAA[f].end(), (u8)0); }
      for (size_t ci = 0; ci < components.size(); ++ci) {
        MovingObject* comp = components[ci].get();
        u8* us[4] = {mask_noAA[0].data(), mask_noAA[1].data(), mask_AA[0].data(), mask_AA[1].data()};
        const u8* vs[4] = {comp->mask_noAA[0].data(), comp->mask_noAA[1].data(), comp->mask_AA[0].data(), comp->mask_AA[1].data()};
        for (int k = 0; k < 4; ++k) {
          u8* u = us[k]; const u8* v = vs[k];
          if (component_modes[ci])
            for (size_t i = 0; i < n; ++i) u[i] = static_cast<u8>(255.f * (1.f - (1.f - u[i] / 255.f) * (1.f - v[i] / 255.f)));
          else
            for (size_t i = 0; i < n; ++i) u[i] = static_cast<u8>(255.f * ((u[i] / 255.f) * (1.f - v[i] / 255.f)));
        }
      }
    }
  }
  void get_point_flow(float* x, float* y, bool inverse = false) const {
    const int W = ctx->W, H = ctx->H;
    if (is_bg) {  // DG.cpp:692-718
      double ix = *x + W / 2, iy = *y + H / 2;
      float save_x = ix, save_y = iy;
      intrinsic_inv.transform(&ix, &iy);
      if (inverse) motion_inv.transform(&ix, &iy);
      else motion.transform(&ix, &iy);
      intrinsic.transform(&ix, &iy);
      *x = ix - save_x; *y = iy - save_y;
      if (has_fields && ix >= 0 && ix < 2 * W && iy >= 0 && iy < 2 * H) {
        *x += cimg_linear_neumann(field, (float)ix, (float)iy, 0);
        *y += cimg_linear_neumann(field, (float)ix, (float)iy, 1);
      }
      return;
    }
    double ix = *x, iy = *y;  // DG.cpp:388-407
    float save_x = ix, save_y = iy;
    if (inverse) motion_inv.transform(&ix, &iy);
    else motion.transform(&ix, &iy);
    *x = ix - save_x; *y = iy - save_y;
    if (has_fields && ix >= 0 && ix < W && iy >= 0 && iy < H) {
      *x += cimg_linear_neumann(field, (float)ix, (float)iy, 0);
      *y += cimg_linear_neumann(field, (float)ix, (float)iy, 1);
    }
  }
};

static Img<u8> pool_texture(const Ctx& c, int raw_index) {  // TextureCollection::getTexturePtr, DG.cpp:158-161
  const oracle_config& g = *c.cfg;
  const size_t idx = (size_t)raw_index % (size_t)g.n_tex;
  if (g.tex_sizes) return Img<u8>::wrap(c.textures + g.tex_offsets[idx], g.tex_sizes[2 * idx], g.tex_sizes[2 * idx + 1], 3);
  return Img<u8>::wrap(c.textures + idx * (size_t)g.tex_w * g.tex_h * 3, g.tex_w, g.tex_h, 3);
}
static void load_field(const Ctx& c, int id, Img<float>& flow, Img<float>& iflow) {
  const int fw = c.W + 1, fh = c.H + 1;
  const size_t plane = (size_t)fw * fh * 2;
  if (id < 0 || id >= c.cfg->n_fields || !c.fields) throw std::runtime_error("oracle: deformed object without an injected field");
  flow = Img<float>(fw, fh, 2); iflow = Img<float>(fw, fh, 2);
  std::memcpy(flow.d.data(), c.fields + (size_t)id * 2 * plane, plane * sizeof(float));
  std::memcpy(iflow.d.data(), c.fields + (size_t)id * 2 * plane + plane, plane * sizeof(float));
}

// RealizeObjectBlueprint, DG.cpp:1065-1173
static std::unique_ptr<MovingObject> realize(const Ctx& c, const ofdg_task_batch& tb, const ofdg_blueprint& p,
                                             const TransAffine& bg_motion, MovingObject* parent) {
  std::unique_ptr<MovingObject> o(new MovingObject(&c, parent ? 0 : (size_t)p.obj_id));
  o->kind = p.obj_type;
  o->is_component = parent != nullptr;
  switch (p.obj_type) {
    case OFDG_OBJ_ELLIPSE:
      o->ellipse.init(0, 0, p.ellipse_scale_x, p.ellipse_scale_y, 100);
      break;
    case OFDG_OBJ_POLYGON: {
      const int32_t* st = tb.seg_type + p.seg_begin;
      const float* sx = tb.seg_x + p.seg_begin; const float* sy = tb.seg_y + p.seg_begin;
      o->path.remove_all();
      o->path.move_to(sx[0], sy[0]);
      for (int i = 1; i < p.seg_count; ++i) {
        switch (st[i]) {
          case OFDG_SEG_LINE: o->path.line_to(sx[i], sy[i]); break;
          case OFDG_SEG_CURVE3: o->path.curve3(sx[i], sy[i], sx[i + 1], sy[i + 1]); ++i; break;
          default: throw std::runtime_error("PolySegmentType_t::Dummy found, this should have been skipped!");
        }
      }
      o->path.close_polygon();
      break;
    }
    case OFDG_OBJ_COMPOSITE: {
      if (c.mode == 9 && p.do_warpfield_deformation) { load_field(c, p.field_id, o->field, o->field_inv); o->has_fields = true; }
      for (int ci = 0; ci < p.comp_count; ++ci) {
        const ofdg_blueprint& cp = tb.blueprints[p.comp_begin + ci];
        o->components.push_back(realize(c, tb, cp, bg_motion, o.get()));
        o->component_modes.push_back(cp.is_additive_component != 0);
      }
      break;
    }
    default:
      throw std::runtime_error("(RealizeObjectBlueprint) Bad object type, or not intended in this mode");
  }
  if (!parent || c.faithful) {  // components crop a texture they never use
    if (c.faithful) o->textures.push_back(randomized_crop(pool_texture(c, p.tex_id), c.W, c.H, 0.f, 1.f, 0, 0));
    else {  // the default-argument chain is exactly the centre crop when the texture is large enough
      Img<u8> t = pool_texture(c, p.tex_id);
      if (t.w < c.W || t.h < c.H) o->textures.push_back(randomized_crop(t, c.W, c.H, 0.f, 1.f, 0, 0));
      else o->textures.push_back(cimg_get_crop_mirror(t, t.w / 2 - c.W / 2, t.h / 2 - c.H / 2, t.w / 2 - c.W / 2 + c.W - 1, t.h / 2 - c.H / 2 + c.H - 1));
    }
  }
  o->set_intrinsic(p.init_rot, p.init_trans_x, p.init_trans_y);
  o->set_motion(p.rot, p.scale, p.trans_x, p.trans_y);
  o->add_background_motion(bg_motion);
  if (c.mode == 9 && p.do_warpfield_deformation) {
    if (parent) { o->field = parent->field; o->field_inv = parent->field_inv; o->has_fields = parent->has_fields; }
    else if (!o->has_fields) { load_field(c, p.field_id, o->field, o->field_inv); o->has_fields = true; }
  }
  // Process_UnfinishedObjectContainer, DG.cpp:726-732 (components first: the composite waits for them)
  o->render_transformed_texture();
  o->render_masks();
  return o;
}

// Process_TaskBucket, DG.cpp:1175-1254
static void process_task(const Ctx& c, const ofdg_task_batch& tb, int t, float* img0, float* img1, float* flow, const oracle_debug* dbg) {
  const int W = c.W, H = c.H;
  const size_t P = (size_t)W * H;
  const int b0 = tb.task_begin[t], b1 = tb.task_begin[t + 1];
  std::map<size_t, std::unique_ptr<MovingObject>> objects;
  const ofdg_blueprint& bp = tb.blueprints[b0];
  std::unique_ptr<MovingObject> bg(new MovingObject(&c, (size_t)bp.obj_id));
  bg->is_bg = true;
  bg->kind = OFDG_OBJ_POLYGON;
  bg->set_intrinsic(0.f, W, H);  // DG.cpp:662
  bg->textures.push_back(randomized_crop(pool_texture(c, bp.tex_id), 2 * W, 2 * H, bp.tex_rot, bp.tex_scale, bp.tex_shift_x, bp.tex_shift_y));
  bg->set_motion(bp.rot, bp.scale, bp.trans_x, bp.trans_y);
  if (c.mode == 9 && bp.do_warpfield_deformation) {  // DG.cpp:1194-1202
    Img<float> f, fi;
    load_field(c, bp.field_id, f, fi);
    f = cimg_get_resize_linear(f, 2 * W, 2 * H);
    fi = cimg_get_resize_linear(fi, 2 * W, 2 * H);
    for (float& v : f.d) v = (float)(v * 2.);
    for (float& v : fi.d) v = (float)(v * 2.);
    bg->field = f; bg->field_inv = fi; bg->has_fields = true;
  }
  bg->render_transformed_texture();
  bg->render_masks();
  const TransAffine bg_motion = bg->motion;
  objects[bg->ID] = std::move(bg);
  for (int bi = b0 + 1; bi < b1; ++bi) {
    const ofdg_blueprint& p = tb.blueprints[bi];
    if (p.parent >= 0) continue;
    std::unique_ptr<MovingObject> o = realize(c, tb, p, bg_motion, nullptr);
    size_t id = o->ID;
    objects[id] = std::move(o);
  }

  // RenderCore::blitObject in ascending-ID order, DG.cpp:762-799, 1216-1223
  Img<u8> frame0(W, H, 3, 0), frame1(W, H, 3, 0);
  std::vector<size_t> index0(P, 0), index1(P, 0);
  int k = 0;
  for (auto& kv : objects) {
    MovingObject& obj = *kv.second;
    for (size_t i = 0; i < P; ++i) if (obj.mask_noAA[0][i] == 255) index0[i] = obj.ID;
    for (size_t i = 0; i < P; ++i) if (obj.mask_noAA[1][i] == 255) index1[i] = obj.ID;
    Img<u8> m0(W, H, 1), m1(W, H, 1);
    m0.d = c.use_aa ? obj.mask_AA[0] : obj.mask_noAA[0];
    m1.d = c.use_aa ? obj.mask_AA[1] : obj.mask_noAA[1];
    cimg_draw_image_masked(frame0, obj.textures[1], m0);
    cimg_draw_image_masked(frame1, obj.textures[2], m1);
    if (dbg && dbg->masks && !obj.is_bg) {
      if (k < dbg->max_objs) {
        u8* dst = dbg->masks + ((size_t)t * dbg->max_objs + k) * 4 * P;
        std::memcpy(dst + 0 * P, obj.mask_AA[0].data(), P);
        std::memcpy(dst + 1 * P, obj.mask_AA[1].data(), P);
        std::memcpy(dst + 2 * P, obj.mask_noAA[0].data(), P);
        std::memcpy(dst + 3 * P, obj.mask_noAA[1].data(), P);
      }
      ++k;
    }
  }
  // RenderCore::computeFlowImage(objects, false), DG.cpp:801-818
  Img<float> flow0(W, H, 2, 0.f);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      size_t idx = index0[(size_t)y * W + x];
      if (idx == 0) continue;
      float xf = x, yf = y;
      objects[idx]->get_point_flow(&xf, &yf);
      flow0.at(x, y, 0) = xf;
      flow0.at(x, y, 1) = yf;
    }
  // copy results, DG.cpp:1229-1245
  for (size_t i = 0; i < 3 * P; ++i) { img0[i] = static_cast<float>(frame0.d[i]); img1[i] = static_cast<float>(frame1.d[i]); }
  if (tb.augment && tb.augment[t].enabled) {  // not in the reference: this repository's augmentation spec (include/ofdg/scene.h)
    const ofdg_augment* a = &tb.augment[t];
    for (int ch = 0; ch < 3; ++ch)
      for (size_t i = 0; i < P; ++i) {
        img0[ch * P + i] = ofdg_augment_value(a, img0[ch * P + i], ch, 0, (uint32_t)i);
        img1[ch * P + i] = ofdg_augment_value(a, img1[ch * P + i], ch, 1, (uint32_t)i);
      }
  }
  std::memcpy(flow, flow0.d.data(), 2 * P * sizeof(float));
  if (dbg && dbg->flow_bw) {  // RenderCore::computeFlowImage(objects, true), DG.cpp:801-818
    float* fb = dbg->flow_bw + (size_t)t * 2 * P;
    std::fill(fb, fb + 2 * P, 0.f);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        size_t idx = index1[(size_t)y * W + x];
        if (idx == 0) continue;
        float xf = x, yf = y;
        objects[idx]->get_point_flow(&xf, &yf, true);
        fb[(size_t)y * W + x] = xf;
        fb[P + (size_t)y * W + x] = yf;
      }
  }
  if (dbg && dbg->occlusion) {  // not in the reference: the product's occlusion top (include/ofdg/ofdg.h)
    float* oc = dbg->occlusion + (size_t)t * P;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t i = (size_t)y * W + x;
        const float tx = (float)x + flow0.at(x, y, 0), ty = (float)y + flow0.at(x, y, 1);
        float occ = 1.f;
        if (tx >= -0.5f && tx < (float)W - 0.5f && ty >= -0.5f && ty < (float)H - 0.5f) {
          const int qx = (int)std::floor(tx + 0.5f), qy = (int)std::floor(ty + 0.5f);
          if (index1[(size_t)qy * W + qx] == index0[i]) occ = 0.f;
        }
        oc[i] = occ;
      }
  }
  if (dbg) {
    if (dbg->id0) for (size_t i = 0; i < P; ++i) dbg->id0[(size_t)t * P + i] = (uint32_t)index0[i];
    if (dbg->id1) for (size_t i = 0; i < P; ++i) dbg->id1[(size_t)t * P + i] = (uint32_t)index1[i];
    if (dbg->frames8) {
      std::memcpy(dbg->frames8 + (size_t)t * 6 * P, frame0.d.data(), 3 * P);
      std::memcpy(dbg->frames8 + (size_t)t * 6 * P + 3 * P, frame1.d.data(), 3 * P);
    }
  }
}

}  // namespace orc

// =================================================================================================
extern "C" {

const char* oracle_last_error(void) { return orc::g_err.c_str(); }

int oracle_render(const oracle_config* cfg, const ofdg_task_batch* tasks, const uint8_t* textures, const float* fields,
                  float* img0, float* img1, float* flow, const oracle_debug* dbg) {
  try {
    orc::Ctx c{cfg->W, cfg->H, cfg->mode, cfg->use_antialiasing != 0, cfg->faithful_copies != 0, cfg, textures, fields};
    const size_t P = (size_t)cfg->W * cfg->H;
    const int n = tasks->n_tasks, nt = std::max(1, std::min(cfg->n_threads, n));
    std::atomic<int> next(0);
    std::atomic<bool> failed(false);
    std::string err;
    auto worker = [&]() {  // WorkerThreadLoop: first-level threads pull tasks from one queue (DG.cpp:1256-1306)
      for (;;) {
        int t = next.fetch_add(1);
        if (t >= n || failed.load()) return;
        try {
          orc::process_task(c, *tasks, t, img0 + (size_t)t * 3 * P, img1 + (size_t)t * 3 * P, flow + (size_t)t * 2 * P, dbg);
        } catch (const std::exception& e) {
          if (!failed.exchange(true)) err = e.what();
        }
      }
    };
    if (nt == 1) worker();
    else {
      std::vector<std::thread> th;
      for (int i = 0; i < nt; ++i) th.emplace_back(worker);
      for (auto& x : th) x.join();
    }
    if (failed.load()) { orc::g_err = err; return 1; }
    return 0;
  } catch (const std::exception& e) {
    orc::g_err = e.what();
    return 1;
  }
}

int oracle_raster_polygon(const double* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask) {
  try {
    orc::Rasterizer ras;
    orc::PathStorage path;
    for (int i = 0; i < n; ++i) { if (i == 0) path.move_to(xy[0], xy[1]); else path.line_to(xy[2 * i], xy[2 * i + 1]); }
    path.close_polygon();
    ras.add_path(path);
    if (aa) ras.set_gamma([](double x) { return x; }); else ras.set_gamma([](double x) { return (x < 0.5) ? 0.0 : 1.0; });
    std::memset(mask, 0, (size_t)W * H);
    orc::render_scanlines_gray8(ras, mask, W, H);
    return 0;
  } catch (const std::exception& e) { orc::g_err = e.what(); return 1; }
}

int oracle_raster_fixed(const int32_t* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask) {
  try {
    orc::Rasterizer ras;
    for (int i = 0; i < n; ++i) { if (i == 0) ras.move_to_fixed(xy[0], xy[1]); else ras.line_to_fixed(xy[2 * i], xy[2 * i + 1]); }
    ras.close_polygon();
    if (aa) ras.set_gamma([](double x) { return x; }); else ras.set_gamma([](double x) { return (x < 0.5) ? 0.0 : 1.0; });
    std::memset(mask, 0, (size_t)W * H);
    orc::render_scanlines_gray8(ras, mask, W, H);
    return 0;
  } catch (const std::exception& e) { orc::g_err = e.what(); return 1; }
}

int oracle_transform_texture(const uint8_t* in, int32_t w, int32_t h, const double* m, uint8_t* out) {
  try {
    orc::Img<orc::u8> src(w, h, 3);
    std::memcpy(src.d.data(), in, src.d.size());
    orc::Img<orc::u8> r = orc::get_transformed_texture(src, orc::TransAffine(m[0], m[1], m[2], m[3], m[4], m[5]));
    std::memcpy(out, r.d.data(), r.d.size());
    return 0;
  } catch (const std::exception& e) { orc::g_err = e.what(); return 1; }
}

int oracle_randomized_crop(const uint8_t* tex, int32_t tw, int32_t th, int32_t out_w, int32_t out_h, float angle, float zoom,
                           int32_t shift_x, int32_t shift_y, uint8_t* out) {
  try {
    orc::Img<orc::u8> src = orc::Img<orc::u8>::wrap(tex, tw, th, 3);
    orc::Img<orc::u8> r = orc::randomized_crop(src, out_w, out_h, angle, zoom, shift_x, shift_y);
    if (r.w != out_w || r.h != out_h) throw std::runtime_error("unexpected crop size");
    std::memcpy(out, r.d.data(), r.d.size());
    return 0;
  } catch (const std::exception& e) { orc::g_err = e.what(); return 1; }
}

void oracle_composite_luts(uint8_t* add_lut, uint8_t* sub_lut) {
  for (int u = 0; u < 256; ++u)
    for (int v = 0; v < 256; ++v) {
      add_lut[u * 256 + v] = static_cast<uint8_t>(255.f * (1.f - (1.f - u / 255.f) * (1.f - v / 255.f)));
      sub_lut[u * 256 + v] = static_cast<uint8_t>(255.f * ((u / 255.f) * (1.f - v / 255.f)));
    }
}

}  // extern "C"
