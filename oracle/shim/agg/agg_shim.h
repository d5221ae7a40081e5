// TEST INFRASTRUCTURE, NOT PRODUCT CODE (oracle/shim).
//
// A from-scratch stand-in for the part of Anti-Grain Geometry 2.4 that the reference generator calls
// (/root/reference/src/caffe/DataGenerator.cpp:24-41 lists the headers, :183-225, :272-277, :304-334,
// :354-362, :462-474, :493-531 are the call sites). The reference does not vendor AGG (it downloads
// agg-2.4.tar.gz, MD5 863d9992fd83c5d40fe1c011501ecf0e, /root/reference/cmake/Dependencies.cmake:4-22) and
// the tarball is not available offline, so this header RESTATES the published AGG 2.4 algorithms behind
// AGG's own class names and call signatures. With it the reference's own, untouched DataGenerator.cpp /
// WarpFields.cpp / data_generation_layer.cpp compile (oracle/ref_build.sh) and every statement the
// reference authors wrote is executed as they wrote it; only the third-party arithmetic below is a
// restatement. PARITY OF THAT PART REMAINS UNPINNED: put the real AGG include directory
// first on the include path (ref_build.sh AGG_INCLUDE=...) to replace it.
//
// Classes follow AGG's public interface; internals are written for brevity (std::vector instead of AGG's
// block allocators, std::sort instead of its quick sort -- cells of equal (y, x) are summed, so order inside
// a run is immaterial).
#ifndef OFDG_ORACLE_AGG_SHIM_H_
#define OFDG_ORACLE_AGG_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace agg {

typedef unsigned char int8u;
typedef short int16;
typedef unsigned short int16u;
typedef int int32;
typedef unsigned int32u;
typedef unsigned char cover_type;

const double pi = 3.14159265358979323846;

inline int iround(double v) { return int((v < 0.0) ? v - 0.5 : v + 0.5); }
inline unsigned uround(double v) { return unsigned(v + 0.5); }

enum cover_scale_e { cover_shift = 8, cover_size = 1 << cover_shift, cover_mask = cover_size - 1, cover_none = 0, cover_full = cover_mask };
enum poly_subpixel_scale_e { poly_subpixel_shift = 8, poly_subpixel_scale = 1 << poly_subpixel_shift, poly_subpixel_mask = poly_subpixel_scale - 1 };
enum image_subpixel_scale_e { image_subpixel_shift = 8, image_subpixel_scale = 1 << image_subpixel_shift, image_subpixel_mask = image_subpixel_scale - 1 };
enum filling_rule_e { fill_non_zero, fill_even_odd };

enum path_commands_e {
  path_cmd_stop = 0, path_cmd_move_to = 1, path_cmd_line_to = 2, path_cmd_curve3 = 3, path_cmd_curve4 = 4,
  path_cmd_curveN = 5, path_cmd_catrom = 6, path_cmd_ubspline = 7, path_cmd_end_poly = 0x0F, path_cmd_mask = 0x0F
};
enum path_flags_e { path_flags_none = 0, path_flags_ccw = 0x10, path_flags_cw = 0x20, path_flags_close = 0x40, path_flags_mask = 0xF0 };

inline bool is_vertex(unsigned c) { return c >= path_cmd_move_to && c < path_cmd_end_poly; }
inline bool is_stop(unsigned c) { return c == path_cmd_stop; }
inline bool is_move_to(unsigned c) { return c == path_cmd_move_to; }
inline bool is_close(unsigned c) { return (c & ~(unsigned)(path_flags_cw | path_flags_ccw)) == (path_cmd_end_poly | path_flags_close); }

struct point_d { double x, y; point_d() {} point_d(double x_, double y_) : x(x_), y(y_) {} };
struct rect_i { int x1, y1, x2, y2; rect_i() {} rect_i(int a, int b, int c, int d) : x1(a), y1(b), x2(c), y2(d) {} };

// ------------------------------------------------------------------------------------------------ trans_affine
struct trans_affine {
  double sx, shy, shx, sy, tx, ty;
  trans_affine() : sx(1.0), shy(0.0), shx(0.0), sy(1.0), tx(0.0), ty(0.0) {}
  trans_affine(double v0, double v1, double v2, double v3, double v4, double v5) : sx(v0), shy(v1), shx(v2), sy(v3), tx(v4), ty(v5) {}
  const trans_affine& multiply(const trans_affine& m) {
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
  double determinant_reciprocal() const { return 1.0 / (sx * sy - shy * shx); }
  const trans_affine& invert() {
    double d = determinant_reciprocal();
    double t0 = sy * d;
    sy = sx * d;
    shy = -shy * d;
    shx = -shx * d;
    double t4 = -tx * t0 - ty * shx;
    ty = -tx * shy - ty * sy;
    sx = t0;
    tx = t4;
    return *this;
  }
  const trans_affine& operator*=(const trans_affine& m) { return multiply(m); }
  trans_affine operator*(const trans_affine& m) const { return trans_affine(*this).multiply(m); }
  void transform(double* x, double* y) const {
    double tmp = *x;
    *x = tmp * sx + *y * shx + tx;
    *y = tmp * shy + *y * sy + ty;
  }
};
struct trans_affine_rotation : trans_affine {
  trans_affine_rotation(double a) : trans_affine(std::cos(a), std::sin(a), -std::sin(a), std::cos(a), 0.0, 0.0) {}
};
struct trans_affine_scaling : trans_affine {
  trans_affine_scaling(double x, double y) : trans_affine(x, 0.0, 0.0, y, 0.0, 0.0) {}
  trans_affine_scaling(double s) : trans_affine(s, 0.0, 0.0, s, 0.0, 0.0) {}
};
struct trans_affine_translation : trans_affine {
  trans_affine_translation(double x, double y) : trans_affine(1.0, 0.0, 0.0, 1.0, x, y) {}
};

// ------------------------------------------------------------------------------------------------ colours
struct rgba8 {
  typedef int8u value_type;
  typedef int32u calc_type;
  enum base_scale_e { base_shift = 8, base_scale = 1 << base_shift, base_mask = base_scale - 1 };
  value_type r, g, b, a;
  rgba8() {}
  rgba8(unsigned r_, unsigned g_, unsigned b_, unsigned a_ = base_mask) : r(value_type(r_)), g(value_type(g_)), b(value_type(b_)), a(value_type(a_)) {}
};
struct gray8 {
  typedef int8u value_type;
  typedef int32u calc_type;
  enum base_scale_e { base_shift = 8, base_scale = 1 << base_shift, base_mask = base_scale - 1 };
  value_type v, a;
  gray8() {}
  gray8(unsigned v_, unsigned a_ = base_mask) : v(int8u(v_)), a(int8u(a_)) {}
};
struct order_rgb { enum rgb_e { R = 0, G = 1, B = 2, rgb_tag }; };

// ------------------------------------------------------------------------------------------------ rendering_buffer
class rendering_buffer {  // row_accessor<int8u>
 public:
  rendering_buffer() : m_buf(0), m_start(0), m_width(0), m_height(0), m_stride(0) {}
  rendering_buffer(int8u* buf, unsigned width, unsigned height, int stride) { attach(buf, width, height, stride); }
  void attach(int8u* buf, unsigned width, unsigned height, int stride) {
    m_buf = m_start = buf;
    m_width = width;
    m_height = height;
    m_stride = stride;
    if (stride < 0) m_start = m_buf - int(height - 1) * stride;
  }
  int8u* buf() { return m_buf; }
  unsigned width() const { return m_width; }
  unsigned height() const { return m_height; }
  int stride() const { return m_stride; }
  int8u* row_ptr(int, int y, unsigned) { return m_start + y * m_stride; }
  int8u* row_ptr(int y) { return m_start + y * m_stride; }
  const int8u* row_ptr(int y) const { return m_start + y * m_stride; }

 private:
  int8u* m_buf;
  int8u* m_start;
  unsigned m_width, m_height;
  int m_stride;
};

// ------------------------------------------------------------------------------------------------ pixel formats
// pixfmt_alpha_blend_gray<blender_gray<gray8>, rendering_buffer, 1, 0>
class pixfmt_gray8 {
 public:
  typedef gray8 color_type;
  typedef int8u value_type;
  typedef int32u calc_type;
  enum { base_shift = 8, base_mask = 255, pix_width = 1 };
  pixfmt_gray8() : m_rbuf(0) {}
  explicit pixfmt_gray8(rendering_buffer& rb) : m_rbuf(&rb) {}
  unsigned width() const { return m_rbuf->width(); }
  unsigned height() const { return m_rbuf->height(); }
  int8u* row_ptr(int y) { return m_rbuf->row_ptr(y); }
  void copy_hline(int x, int y, unsigned len, const color_type& c) {
    value_type* p = m_rbuf->row_ptr(x, y, len) + x;
    do { *p = c.v; ++p; } while (--len);
  }
  void blend_solid_hspan(int x, int y, unsigned len, const color_type& c, const int8u* covers) {
    if (c.a) {
      value_type* p = m_rbuf->row_ptr(x, y, len) + x;
      do {
        calc_type alpha = (calc_type(c.a) * (calc_type(*covers) + 1)) >> 8;
        if (alpha == base_mask) *p = c.v;
        else *p = (value_type)((((calc_type(c.v) - calc_type(*p)) * alpha) + (calc_type(*p) << base_shift)) >> base_shift);  // blender_gray::blend_pix
        ++p;
        ++covers;
      } while (--len);
    }
  }
  void blend_hline(int x, int y, unsigned len, const color_type& c, int8u cover) {
    if (c.a) {
      value_type* p = m_rbuf->row_ptr(x, y, len) + x;
      calc_type alpha = (calc_type(c.a) * (calc_type(cover) + 1)) >> 8;
      if (alpha == base_mask) { do { *p = c.v; ++p; } while (--len); }
      else do { *p = (value_type)((((calc_type(c.v) - calc_type(*p)) * alpha) + (calc_type(*p) << base_shift)) >> base_shift); ++p; } while (--len);
    }
  }

 private:
  rendering_buffer* m_rbuf;
};

// pixfmt_alpha_blend_rgb<blender_rgb<rgba8, order_rgb>, rendering_buffer>
class pixfmt_rgb24 {
 public:
  typedef rgba8 color_type;
  typedef order_rgb order_type;
  typedef int8u value_type;
  typedef int32u calc_type;
  enum { base_shift = 8, base_mask = 255, pix_width = 3 };
  pixfmt_rgb24() : m_rbuf(0) {}
  explicit pixfmt_rgb24(rendering_buffer& rb) : m_rbuf(&rb) {}
  unsigned width() const { return m_rbuf->width(); }
  unsigned height() const { return m_rbuf->height(); }
  int8u* row_ptr(int y) { return m_rbuf->row_ptr(y); }
  const int8u* row_ptr(int y) const { return m_rbuf->row_ptr(y); }
  void copy_hline(int x, int y, unsigned len, const color_type& c) {
    value_type* p = m_rbuf->row_ptr(x, y, len) + x + x + x;
    do { p[0] = c.r; p[1] = c.g; p[2] = c.b; p += 3; } while (--len);
  }
  void blend_color_hspan(int x, int y, unsigned len, const color_type* colors, const int8u* covers, int8u cover) {
    value_type* p = m_rbuf->row_ptr(x, y, len) + x + x + x;
    do {
      copy_or_blend_pix(p, *colors++, covers ? unsigned(*covers++) : unsigned(cover));
      p += 3;
    } while (--len);
  }

 private:
  void copy_or_blend_pix(value_type* p, const color_type& c, unsigned cover) {
    if (c.a) {
      calc_type alpha = (calc_type(c.a) * (cover + 1)) >> 8;
      if (alpha == base_mask) { p[0] = c.r; p[1] = c.g; p[2] = c.b; }
      else {  // blender_rgb::blend_pix
        p[0] += (value_type)(((calc_type(c.r) - calc_type(p[0])) * alpha) >> base_shift);
        p[1] += (value_type)(((calc_type(c.g) - calc_type(p[1])) * alpha) >> base_shift);
        p[2] += (value_type)(((calc_type(c.b) - calc_type(p[2])) * alpha) >> base_shift);
      }
    }
  }
  rendering_buffer* m_rbuf;
};

// ------------------------------------------------------------------------------------------------ renderer_base
template <class PixelFormat>
class renderer_base {
 public:
  typedef PixelFormat pixfmt_type;
  typedef typename pixfmt_type::color_type color_type;
  renderer_base() : m_ren(0), m_clip_box(1, 1, 0, 0) {}
  explicit renderer_base(pixfmt_type& ren) : m_ren(&ren), m_clip_box(0, 0, ren.width() - 1, ren.height() - 1) {}
  unsigned width() const { return m_ren->width(); }
  unsigned height() const { return m_ren->height(); }
  int xmin() const { return m_clip_box.x1; }
  int ymin() const { return m_clip_box.y1; }
  int xmax() const { return m_clip_box.x2; }
  int ymax() const { return m_clip_box.y2; }
  void clear(const color_type& c) {
    if (width())
      for (unsigned y = 0; y < height(); y++) m_ren->copy_hline(0, y, width(), c);
  }
  void blend_hline(int x1, int y, int x2, const color_type& c, cover_type cover) {
    if (x1 > x2) { int t = x2; x2 = x1; x1 = t; }
    if (y > ymax()) return;
    if (y < ymin()) return;
    if (x1 > xmax()) return;
    if (x2 < xmin()) return;
    if (x1 < xmin()) x1 = xmin();
    if (x2 > xmax()) x2 = xmax();
    m_ren->blend_hline(x1, y, x2 - x1 + 1, c, cover);
  }
  void blend_solid_hspan(int x, int y, int len, const color_type& c, const cover_type* covers) {
    if (y > ymax()) return;
    if (y < ymin()) return;
    if (x < xmin()) {
      len -= xmin() - x;
      if (len <= 0) return;
      covers += xmin() - x;
      x = xmin();
    }
    if (x + len > xmax()) {
      len = xmax() - x + 1;
      if (len <= 0) return;
    }
    m_ren->blend_solid_hspan(x, y, len, c, covers);
  }
  void blend_color_hspan(int x, int y, int len, const color_type* colors, const cover_type* covers, cover_type cover = cover_full) {
    if (y > ymax()) return;
    if (y < ymin()) return;
    if (x < xmin()) {
      int d = xmin() - x;
      len -= d;
      if (len <= 0) return;
      if (covers) covers += d;
      colors += d;
      x = xmin();
    }
    if (x + len > xmax()) {
      len = xmax() - x + 1;
      if (len <= 0) return;
    }
    m_ren->blend_color_hspan(x, y, len, colors, covers, cover);
  }

 private:
  pixfmt_type* m_ren;
  rect_i m_clip_box;
};

// ------------------------------------------------------------------------------------------------ scanline_u8
class scanline_u8 {
 public:
  typedef int8u cover_type;
  typedef int16 coord_type;
  struct span { coord_type x; coord_type len; cover_type* covers; };
  typedef span* iterator;
  typedef const span* const_iterator;
  scanline_u8() : m_min_x(0), m_last_x(0x7FFFFFF0), m_y(0), m_cur_span(0) {}
  void reset(int min_x, int max_x) {
    unsigned max_len = max_x - min_x + 2;
    if (max_len > m_spans.size()) { m_spans.resize(max_len); m_covers.resize(max_len); }
    m_last_x = 0x7FFFFFF0;
    m_min_x = min_x;
    m_cur_span = &m_spans[0];
  }
  void add_cell(int x, unsigned cover) {
    x -= m_min_x;
    m_covers[x] = (cover_type)cover;
    if (x == m_last_x + 1) m_cur_span->len++;
    else {
      m_cur_span++;
      m_cur_span->x = (coord_type)(x + m_min_x);
      m_cur_span->len = 1;
      m_cur_span->covers = &m_covers[x];
    }
    m_last_x = x;
  }
  void add_span(int x, unsigned len, unsigned cover) {
    x -= m_min_x;
    std::memset(&m_covers[x], cover, len);
    if (x == m_last_x + 1) m_cur_span->len += (coord_type)len;
    else {
      m_cur_span++;
      m_cur_span->x = (coord_type)(x + m_min_x);
      m_cur_span->len = (coord_type)len;
      m_cur_span->covers = &m_covers[x];
    }
    m_last_x = x + len - 1;
  }
  void finalize(int y) { m_y = y; }
  void reset_spans() { m_last_x = 0x7FFFFFF0; m_cur_span = &m_spans[0]; }
  int y() const { return m_y; }
  unsigned num_spans() const { return unsigned(m_cur_span - &m_spans[0]); }
  const_iterator begin() const { return &m_spans[1]; }

 private:
  int m_min_x, m_last_x, m_y;
  std::vector<cover_type> m_covers;
  std::vector<span> m_spans;
  span* m_cur_span;
};

// ------------------------------------------------------------------------------------------------ rasterizer
struct cell_aa { int x, y, cover, area; };

class rasterizer_cells_aa {
 public:
  rasterizer_cells_aa() { reset(); }
  void reset() {
    m_cells.clear();
    m_curr.x = 0x7FFFFFFF; m_curr.y = 0x7FFFFFFF; m_curr.cover = 0; m_curr.area = 0;
    m_sorted = false;
    m_min_x = 0x7FFFFFFF; m_min_y = 0x7FFFFFFF; m_max_x = -0x7FFFFFFF; m_max_y = -0x7FFFFFFF;
  }
  int min_x() const { return m_min_x; }
  int min_y() const { return m_min_y; }
  int max_x() const { return m_max_x; }
  int max_y() const { return m_max_y; }
  bool sorted() const { return m_sorted; }
  unsigned total_cells() const { return (unsigned)m_cells.size(); }
  const std::vector<cell_aa>& cells() const { return m_cells; }

  void line(int x1, int y1, int x2, int y2) {
    enum dx_limit_e { dx_limit = 16384 << poly_subpixel_shift };
    int dx = x2 - x1;
    if (dx >= dx_limit || dx <= -dx_limit) {
      int cx = (x1 + x2) >> 1;
      int cy = (y1 + y2) >> 1;
      line(x1, y1, cx, cy);
      line(cx, cy, x2, y2);
      return;  // (never reached by the reference: no edge spans 16384 pixels)
    }
    int dy = y2 - y1;
    int ex1 = x1 >> poly_subpixel_shift, ex2 = x2 >> poly_subpixel_shift;
    int ey1 = y1 >> poly_subpixel_shift, ey2 = y2 >> poly_subpixel_shift;
    int fy1 = y1 & poly_subpixel_mask, fy2 = y2 & poly_subpixel_mask;
    int x_from, x_to, p, rem, mod, lift, delta, first, incr;
    if (ex1 < m_min_x) m_min_x = ex1;
    if (ex1 > m_max_x) m_max_x = ex1;
    if (ey1 < m_min_y) m_min_y = ey1;
    if (ey1 > m_max_y) m_max_y = ey1;
    if (ex2 < m_min_x) m_min_x = ex2;
    if (ex2 > m_max_x) m_max_x = ex2;
    if (ey2 < m_min_y) m_min_y = ey2;
    if (ey2 > m_max_y) m_max_y = ey2;
    set_curr_cell(ex1, ey1);
    if (ey1 == ey2) {  // everything is on a single hline
      render_hline(ey1, x1, fy1, x2, fy2);
      return;
    }
    incr = 1;
    if (dx == 0) {  // vertical line: only one cell per row
      int ex = x1 >> poly_subpixel_shift;
      int two_fx = (x1 - (ex << poly_subpixel_shift)) << 1;
      int area;
      first = poly_subpixel_scale;
      if (dy < 0) { first = 0; incr = -1; }
      x_from = x1;
      delta = first - fy1;
      m_curr.cover += delta;
      m_curr.area += two_fx * delta;
      ey1 += incr;
      set_curr_cell(ex, ey1);
      delta = first + first - poly_subpixel_scale;
      area = two_fx * delta;
      while (ey1 != ey2) {
        m_curr.cover = delta;
        m_curr.area = area;
        ey1 += incr;
        set_curr_cell(ex, ey1);
      }
      delta = fy2 - poly_subpixel_scale + first;
      m_curr.cover += delta;
      m_curr.area += two_fx * delta;
      return;
    }
    // several hlines
    p = (poly_subpixel_scale - fy1) * dx;
    first = poly_subpixel_scale;
    if (dy < 0) { p = fy1 * dx; first = 0; incr = -1; dy = -dy; }
    delta = p / dy;
    mod = p % dy;
    if (mod < 0) { delta--; mod += dy; }
    x_from = x1 + delta;
    render_hline(ey1, x1, fy1, x_from, first);
    ey1 += incr;
    set_curr_cell(x_from >> poly_subpixel_shift, ey1);
    if (ey1 != ey2) {
      p = poly_subpixel_scale * dx;
      lift = p / dy;
      rem = p % dy;
      if (rem < 0) { lift--; rem += dy; }
      mod -= dy;
      while (ey1 != ey2) {
        delta = lift;
        mod += rem;
        if (mod >= 0) { mod -= dy; delta++; }
        x_to = x_from + delta;
        render_hline(ey1, x_from, poly_subpixel_scale - first, x_to, first);
        x_from = x_to;
        ey1 += incr;
        set_curr_cell(x_from >> poly_subpixel_shift, ey1);
      }
    }
    render_hline(ey1, x_from, poly_subpixel_scale - first, x2, fy2);
  }

  void sort_cells() {
    if (m_sorted) return;
    add_curr_cell();
    m_curr.x = 0x7FFFFFFF; m_curr.y = 0x7FFFFFFF; m_curr.cover = 0; m_curr.area = 0;
    if (m_cells.empty()) return;
    std::stable_sort(m_cells.begin(), m_cells.end(), [](const cell_aa& a, const cell_aa& b) { return a.y != b.y ? a.y < b.y : a.x < b.x; });
    m_sorted = true;
  }

 private:
  void add_curr_cell() { if (m_curr.area | m_curr.cover) m_cells.push_back(m_curr); }
  void set_curr_cell(int x, int y) {
    if (m_curr.x != x || m_curr.y != y) {
      add_curr_cell();
      m_curr.x = x; m_curr.y = y; m_curr.cover = 0; m_curr.area = 0;
    }
  }
  void render_hline(int ey, int x1, int y1, int x2, int y2) {
    int ex1 = x1 >> poly_subpixel_shift, ex2 = x2 >> poly_subpixel_shift;
    int fx1 = x1 & poly_subpixel_mask, fx2 = x2 & poly_subpixel_mask;
    int delta, p, first, dx, incr, lift, mod, rem;
    if (y1 == y2) {  // trivial case; happens often
      set_curr_cell(ex2, ey);
      return;
    }
    if (ex1 == ex2) {  // everything is located in a single cell
      delta = y2 - y1;
      m_curr.cover += delta;
      m_curr.area += (fx1 + fx2) * delta;
      return;
    }
    // a run of adjacent cells on the same hline
    p = (poly_subpixel_scale - fx1) * (y2 - y1);
    first = poly_subpixel_scale;
    incr = 1;
    dx = x2 - x1;
    if (dx < 0) { p = fx1 * (y2 - y1); first = 0; incr = -1; dx = -dx; }
    delta = p / dx;
    mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    m_curr.cover += delta;
    m_curr.area += (fx1 + first) * delta;
    ex1 += incr;
    set_curr_cell(ex1, ey);
    y1 += delta;
    if (ex1 != ex2) {
      p = poly_subpixel_scale * (y2 - y1 + delta);
      lift = p / dx;
      rem = p % dx;
      if (rem < 0) { lift--; rem += dx; }
      mod -= dx;
      while (ex1 != ex2) {
        delta = lift;
        mod += rem;
        if (mod >= 0) { mod -= dx; delta++; }
        m_curr.cover += delta;
        m_curr.area += poly_subpixel_scale * delta;
        y1 += delta;
        ex1 += incr;
        set_curr_cell(ex1, ey);
      }
    }
    delta = y2 - y1;
    m_curr.cover += delta;
    m_curr.area += (fx2 + poly_subpixel_scale - first) * delta;
  }

  std::vector<cell_aa> m_cells;
  cell_aa m_curr;
  bool m_sorted;
  int m_min_x, m_min_y, m_max_x, m_max_y;
};

struct gamma_none { double operator()(double x) const { return x; } };
class gamma_threshold {
 public:
  gamma_threshold() : m_threshold(0.5) {}
  gamma_threshold(double t) : m_threshold(t) {}
  double operator()(double x) const { return (x < m_threshold) ? 0.0 : 1.0; }
 private:
  double m_threshold;
};

struct rasterizer_sl_clip_int {};  // the reference never sets a clip box: the clipper is a pass-through

template <class Clip = rasterizer_sl_clip_int>
class rasterizer_scanline_aa {
  enum status { status_initial, status_move_to, status_line_to, status_closed };

 public:
  enum aa_scale_e { aa_shift = 8, aa_scale = 1 << aa_shift, aa_mask = aa_scale - 1, aa_scale2 = aa_scale * 2, aa_mask2 = aa_scale2 - 1 };
  rasterizer_scanline_aa() : m_filling_rule(fill_non_zero), m_auto_close(true), m_start_x(0), m_start_y(0), m_x1(0), m_y1(0), m_status(status_initial), m_scan_y(0), m_sweep(0) {
    for (int i = 0; i < aa_scale; i++) m_gamma[i] = i;
  }
  void reset() { m_outline.reset(); m_status = status_initial; }
  template <class GammaF> void gamma(const GammaF& gamma_function) {
    for (int i = 0; i < aa_scale; i++) m_gamma[i] = uround(gamma_function(double(i) / aa_mask) * aa_mask);
  }
  void close_polygon() {
    if (m_status == status_line_to) {
      clip_line_to(m_start_x, m_start_y);
      m_status = status_closed;
    }
  }
  void move_to_d(double x, double y) {
    if (m_outline.sorted()) reset();
    if (m_auto_close) close_polygon();
    m_x1 = m_start_x = iround(x * poly_subpixel_scale);  // ras_conv_int::upscale; clipper.move_to
    m_y1 = m_start_y = iround(y * poly_subpixel_scale);
    m_status = status_move_to;
  }
  void line_to_d(double x, double y) {
    clip_line_to(iround(x * poly_subpixel_scale), iround(y * poly_subpixel_scale));
    m_status = status_line_to;
  }
  void add_vertex(double x, double y, unsigned cmd) {
    if (is_move_to(cmd)) move_to_d(x, y);
    else if (is_vertex(cmd)) line_to_d(x, y);
    else if (is_close(cmd)) close_polygon();
  }
  template <class VertexSource> void add_path(VertexSource& vs, unsigned path_id = 0) {
    double x, y;
    unsigned cmd;
    vs.rewind(path_id);
    if (m_outline.sorted()) reset();
    while (!is_stop(cmd = vs.vertex(&x, &y))) add_vertex(x, y, cmd);
  }
  int min_x() const { return m_outline.min_x(); }
  int min_y() const { return m_outline.min_y(); }
  int max_x() const { return m_outline.max_x(); }
  int max_y() const { return m_outline.max_y(); }
  bool rewind_scanlines() {
    if (m_auto_close) close_polygon();
    m_outline.sort_cells();
    if (m_outline.total_cells() == 0) return false;
    m_scan_y = m_outline.min_y();
    m_sweep = 0;
    return true;
  }
  unsigned calculate_alpha(int area) const {
    int cover = area >> (poly_subpixel_shift * 2 + 1 - aa_shift);
    if (cover < 0) cover = -cover;
    if (m_filling_rule == fill_even_odd) {
      cover &= aa_mask2;
      if (cover > aa_scale) cover = aa_scale2 - cover;
    }
    if (cover > aa_mask) cover = aa_mask;
    return m_gamma[cover];
  }
  template <class Scanline> bool sweep_scanline(Scanline& sl) {
    const std::vector<cell_aa>& cells = m_outline.cells();
    for (;;) {
      if (m_scan_y > m_outline.max_y()) return false;
      sl.reset_spans();
      size_t j = m_sweep;
      while (j < cells.size() && cells[j].y == m_scan_y) ++j;
      unsigned num_cells = unsigned(j - m_sweep);
      const cell_aa* cp = cells.data() + m_sweep;
      m_sweep = j;
      int cover = 0;
      while (num_cells) {
        const cell_aa* cur_cell = cp;
        int x = cur_cell->x;
        int area = cur_cell->area;
        unsigned alpha;
        cover += cur_cell->cover;
        // accumulate all cells with the same X
        while (--num_cells) {
          cur_cell = ++cp;
          if (cur_cell->x != x) break;
          area += cur_cell->area;
          cover += cur_cell->cover;
        }
        if (area) {
          alpha = calculate_alpha((cover << (poly_subpixel_shift + 1)) - area);
          if (alpha) sl.add_cell(x, alpha);
          x++;
        }
        if (num_cells && cur_cell->x > x) {
          alpha = calculate_alpha(cover << (poly_subpixel_shift + 1));
          if (alpha) sl.add_span(x, cur_cell->x - x, alpha);
        }
      }
      if (sl.num_spans()) break;
      ++m_scan_y;
    }
    sl.finalize(m_scan_y);
    ++m_scan_y;
    return true;
  }

 private:
  void clip_line_to(int x2, int y2) {  // rasterizer_sl_clip<ras_conv_int>::line_to with clipping off
    m_outline.line(m_x1, m_y1, x2, y2);
    m_x1 = x2;
    m_y1 = y2;
  }
  rasterizer_cells_aa m_outline;
  int m_gamma[aa_scale];
  filling_rule_e m_filling_rule;
  bool m_auto_close;
  int m_start_x, m_start_y, m_x1, m_y1;
  unsigned m_status;
  int m_scan_y;
  size_t m_sweep;
};

// ------------------------------------------------------------------------------------------------ scanline renderers
template <class Scanline, class BaseRenderer, class ColorT>
void render_scanline_aa_solid(const Scanline& sl, BaseRenderer& ren, const ColorT& color) {
  int y = sl.y();
  unsigned num_spans = sl.num_spans();
  typename Scanline::const_iterator span = sl.begin();
  for (;;) {
    int x = span->x;
    if (span->len > 0) ren.blend_solid_hspan(x, y, (unsigned)span->len, color, span->covers);
    else ren.blend_hline(x, y, (unsigned)(x - span->len - 1), color, *(span->covers));
    if (--num_spans == 0) break;
    ++span;
  }
}

template <class BaseRenderer>
class renderer_scanline_aa_solid {
 public:
  typedef BaseRenderer base_ren_type;
  typedef typename base_ren_type::color_type color_type;
  renderer_scanline_aa_solid() : m_ren(0) {}
  explicit renderer_scanline_aa_solid(base_ren_type& ren) : m_ren(&ren) {}
  void attach(base_ren_type& ren) { m_ren = &ren; }
  void color(const color_type& c) { m_color = c; }
  const color_type& color() const { return m_color; }
  void prepare() {}
  template <class Scanline> void render(const Scanline& sl) { render_scanline_aa_solid(sl, *m_ren, m_color); }

 private:
  base_ren_type* m_ren;
  color_type m_color;
};

template <class Rasterizer, class Scanline, class Renderer>
void render_scanlines(Rasterizer& ras, Scanline& sl, Renderer& ren) {
  if (ras.rewind_scanlines()) {
    sl.reset(ras.min_x(), ras.max_x());
    ren.prepare();
    while (ras.sweep_scanline(sl)) ren.render(sl);
  }
}

template <class ColorT>
class span_allocator {
 public:
  typedef ColorT color_type;
  color_type* allocate(unsigned span_len) {
    if (span_len > m_span.size()) m_span.resize(((span_len + 255) >> 8) << 8);
    return &m_span[0];
  }
 private:
  std::vector<color_type> m_span;
};

template <class Scanline, class BaseRenderer, class SpanAllocator, class SpanGenerator>
void render_scanline_aa(const Scanline& sl, BaseRenderer& ren, SpanAllocator& alloc, SpanGenerator& span_gen) {
  int y = sl.y();
  unsigned num_spans = sl.num_spans();
  typename Scanline::const_iterator span = sl.begin();
  for (;;) {
    int x = span->x;
    int len = span->len;
    const typename Scanline::cover_type* covers = span->covers;
    if (len < 0) len = -len;
    typename BaseRenderer::color_type* colors = alloc.allocate(len);
    span_gen.generate(colors, x, y, len);
    ren.blend_color_hspan(x, y, len, colors, (span->len < 0) ? 0 : covers, *covers);
    if (--num_spans == 0) break;
    ++span;
  }
}

template <class Rasterizer, class Scanline, class BaseRenderer, class SpanAllocator, class SpanGenerator>
void render_scanlines_aa(Rasterizer& ras, Scanline& sl, BaseRenderer& ren, SpanAllocator& alloc, SpanGenerator& span_gen) {
  if (ras.rewind_scanlines()) {
    sl.reset(ras.min_x(), ras.max_x());
    span_gen.prepare();
    while (ras.sweep_scanline(sl)) render_scanline_aa(sl, ren, alloc, span_gen);
  }
}

// ------------------------------------------------------------------------------------------------ image source + span filter
class wrap_mode_reflect {
 public:
  wrap_mode_reflect() {}
  wrap_mode_reflect(unsigned size) : m_size(size), m_size2(size * 2), m_add(m_size2 * (0x3FFFFFFF / m_size2)), m_value(0) {}
  unsigned operator()(int v) {
    m_value = (unsigned(v) + m_add) % m_size2;
    if (m_value >= m_size) return m_size2 - m_value - 1;
    return m_value;
  }
  unsigned operator++() {
    ++m_value;
    if (m_value >= m_size2) m_value = 0;
    if (m_value >= m_size) return m_size2 - m_value - 1;
    return m_value;
  }
 private:
  unsigned m_size, m_size2, m_add, m_value;
};

template <class PixFmt, class WrapX, class WrapY>
class image_accessor_wrap {
 public:
  typedef PixFmt pixfmt_type;
  typedef typename pixfmt_type::color_type color_type;
  typedef typename pixfmt_type::order_type order_type;
  typedef typename pixfmt_type::value_type value_type;
  enum pix_width_e { pix_width = pixfmt_type::pix_width };
  image_accessor_wrap() {}
  explicit image_accessor_wrap(const pixfmt_type& pixf) : m_pixf(&pixf), m_wrap_x(pixf.width()), m_wrap_y(pixf.height()) {}
  const int8u* span(int x, int y, unsigned) {
    m_x = x;
    m_row_ptr = m_pixf->row_ptr(m_wrap_y(y));
    return m_row_ptr + m_wrap_x(x) * pix_width;
  }
  const int8u* next_x() {
    int x = ++m_wrap_x;
    return m_row_ptr + x * pix_width;
  }
  const int8u* next_y() {
    m_row_ptr = m_pixf->row_ptr(++m_wrap_y);
    return m_row_ptr + m_wrap_x(m_x) * pix_width;
  }
 private:
  const pixfmt_type* m_pixf;
  const int8u* m_row_ptr;
  int m_x;
  WrapX m_wrap_x;
  WrapY m_wrap_y;
};

class dda2_line_interpolator {
 public:
  dda2_line_interpolator() {}
  dda2_line_interpolator(int y1, int y2, int count)
      : m_cnt(count <= 0 ? 1 : count), m_lft((y2 - y1) / m_cnt), m_rem((y2 - y1) % m_cnt), m_mod(m_rem), m_y(y1) {
    if (m_mod <= 0) { m_mod += count; m_rem += count; m_lft--; }
    m_mod -= count;
  }
  void operator++() {
    m_mod += m_rem;
    m_y += m_lft;
    if (m_mod > 0) { m_mod -= m_cnt; m_y++; }
  }
  int y() const { return m_y; }
 private:
  int m_cnt, m_lft, m_rem, m_mod, m_y;
};

template <class Transformer = trans_affine, unsigned SubpixelShift = 8>
class span_interpolator_linear {
 public:
  typedef Transformer trans_type;
  enum subpixel_scale_e { subpixel_shift = SubpixelShift, subpixel_scale = 1 << subpixel_shift };
  span_interpolator_linear() {}
  span_interpolator_linear(const trans_type& trans) : m_trans(&trans) {}
  void begin(double x, double y, unsigned len) {
    double tx, ty;
    tx = x;
    ty = y;
    m_trans->transform(&tx, &ty);
    int x1 = iround(tx * subpixel_scale);
    int y1 = iround(ty * subpixel_scale);
    tx = x + len;
    ty = y;
    m_trans->transform(&tx, &ty);
    int x2 = iround(tx * subpixel_scale);
    int y2 = iround(ty * subpixel_scale);
    m_li_x = dda2_line_interpolator(x1, x2, len);
    m_li_y = dda2_line_interpolator(y1, y2, len);
  }
  void operator++() { ++m_li_x; ++m_li_y; }
  void coordinates(int* x, int* y) const { *x = m_li_x.y(); *y = m_li_y.y(); }
 private:
  const trans_type* m_trans;
  dda2_line_interpolator m_li_x, m_li_y;
};

template <class Source, class Interpolator>
class span_image_filter_rgb_bilinear {
 public:
  typedef Source source_type;
  typedef typename source_type::color_type color_type;
  typedef typename source_type::order_type order_type;
  typedef Interpolator interpolator_type;
  typedef typename color_type::value_type value_type;
  typedef typename color_type::calc_type calc_type;
  enum base_scale_e { base_shift = color_type::base_shift, base_mask = color_type::base_mask };
  span_image_filter_rgb_bilinear() {}
  span_image_filter_rgb_bilinear(source_type& src, interpolator_type& inter) : m_src(&src), m_interpolator(&inter) {}
  void prepare() {}
  void generate(color_type* span, int x, int y, unsigned len) {
    m_interpolator->begin(x + 0.5, y + 0.5, len);  // filter_dx_dbl / filter_dy_dbl
    calc_type fg[3];
    const value_type* fg_ptr;
    do {
      int x_hr, y_hr;
      m_interpolator->coordinates(&x_hr, &y_hr);
      x_hr -= image_subpixel_scale / 2;  // filter_dx_int
      y_hr -= image_subpixel_scale / 2;
      int x_lr = x_hr >> image_subpixel_shift;
      int y_lr = y_hr >> image_subpixel_shift;
      unsigned weight;
      fg[0] = fg[1] = fg[2] = image_subpixel_scale * image_subpixel_scale / 2;
      x_hr &= image_subpixel_mask;
      y_hr &= image_subpixel_mask;
      fg_ptr = (const value_type*)m_src->span(x_lr, y_lr, 2);
      weight = (image_subpixel_scale - x_hr) * (image_subpixel_scale - y_hr);
      fg[0] += weight * *fg_ptr++; fg[1] += weight * *fg_ptr++; fg[2] += weight * *fg_ptr;
      fg_ptr = (const value_type*)m_src->next_x();
      weight = x_hr * (image_subpixel_scale - y_hr);
      fg[0] += weight * *fg_ptr++; fg[1] += weight * *fg_ptr++; fg[2] += weight * *fg_ptr;
      fg_ptr = (const value_type*)m_src->next_y();
      weight = (image_subpixel_scale - x_hr) * y_hr;
      fg[0] += weight * *fg_ptr++; fg[1] += weight * *fg_ptr++; fg[2] += weight * *fg_ptr;
      fg_ptr = (const value_type*)m_src->next_x();
      weight = x_hr * y_hr;
      fg[0] += weight * *fg_ptr++; fg[1] += weight * *fg_ptr++; fg[2] += weight * *fg_ptr;
      span->r = value_type(fg[order_type::R] >> (image_subpixel_shift * 2));
      span->g = value_type(fg[order_type::G] >> (image_subpixel_shift * 2));
      span->b = value_type(fg[order_type::B] >> (image_subpixel_shift * 2));
      span->a = base_mask;
      ++span;
      ++(*m_interpolator);
    } while (--len);
  }
 private:
  source_type* m_src;
  interpolator_type* m_interpolator;
};

// ------------------------------------------------------------------------------------------------ vertex sources
class ellipse {
 public:
  ellipse() : m_x(0.0), m_y(0.0), m_rx(1.0), m_ry(1.0), m_scale(1.0), m_num(4), m_step(0), m_cw(false) {}
  ellipse(double x, double y, double rx, double ry, unsigned num_steps = 0, bool cw = false)
      : m_x(x), m_y(y), m_rx(rx), m_ry(ry), m_scale(1.0), m_num(num_steps), m_step(0), m_cw(cw) { if (m_num == 0) calc_num_steps(); }
  void init(double x, double y, double rx, double ry, unsigned num_steps = 0, bool cw = false) {
    m_x = x; m_y = y; m_rx = rx; m_ry = ry; m_num = num_steps; m_step = 0; m_cw = cw;
    if (m_num == 0) calc_num_steps();
  }
  void rewind(unsigned) { m_step = 0; }
  unsigned vertex(double* x, double* y) {
    if (m_step == m_num) { ++m_step; return path_cmd_end_poly | path_flags_close | path_flags_ccw; }
    if (m_step > m_num) return path_cmd_stop;
    double angle = double(m_step) / double(m_num) * 2.0 * pi;
    if (m_cw) angle = 2.0 * pi - angle;
    *x = m_x + std::cos(angle) * m_rx;
    *y = m_y + std::sin(angle) * m_ry;
    m_step++;
    return ((m_step == 1) ? path_cmd_move_to : path_cmd_line_to);
  }
 private:
  void calc_num_steps() {
    double ra = (std::fabs(m_rx) + std::fabs(m_ry)) / 2;
    double da = std::acos(ra / (ra + 0.125 / m_scale)) * 2;
    m_num = uround(2 * pi / da);
  }
  double m_x, m_y, m_rx, m_ry, m_scale;
  unsigned m_num, m_step;
  bool m_cw;
};

class path_storage {  // path_base<vertex_block_storage<double>>, the calls the reference makes
 public:
  path_storage() : m_iterator(0) {}
  void remove_all() { m_v.clear(); m_iterator = 0; }
  void move_to(double x, double y) { add(x, y, path_cmd_move_to); }
  void line_to(double x, double y) { add(x, y, path_cmd_line_to); }
  void curve3(double x_ctrl, double y_ctrl, double x_to, double y_to) { add(x_ctrl, y_ctrl, path_cmd_curve3); add(x_to, y_to, path_cmd_curve3); }
  void end_poly(unsigned flags = path_flags_close) { if (!m_v.empty() && is_vertex(m_v.back().cmd)) add(0.0, 0.0, path_cmd_end_poly | flags); }
  void close_polygon(unsigned flags = path_flags_none) { end_poly(path_flags_close | flags); }
  unsigned total_vertices() const { return (unsigned)m_v.size(); }
  void rewind(unsigned path_id) { m_iterator = path_id; }
  unsigned vertex(double* x, double* y) {
    if (m_iterator >= m_v.size()) return path_cmd_stop;
    *x = m_v[m_iterator].x;
    *y = m_v[m_iterator].y;
    return m_v[m_iterator++].cmd;
  }
 private:
  struct V { double x, y; unsigned cmd; };
  void add(double x, double y, unsigned cmd) { V v = {x, y, cmd}; m_v.push_back(v); }
  std::vector<V> m_v;
  unsigned m_iterator;
};

template <class VertexSource, class Transformer = trans_affine>
class conv_transform {
 public:
  conv_transform(VertexSource& source, const Transformer& tr) : m_source(&source), m_trans(&tr) {}
  void attach(VertexSource& source) { m_source = &source; }
  void rewind(unsigned path_id) { m_source->rewind(path_id); }
  unsigned vertex(double* x, double* y) {
    unsigned cmd = m_source->vertex(x, y);
    if (is_vertex(cmd)) m_trans->transform(x, y);
    return cmd;
  }
 private:
  VertexSource* m_source;
  const Transformer* m_trans;
};

const double curve_collinearity_epsilon = 1e-30;
const double curve_angle_tolerance_epsilon = 0.01;
enum curve_recursion_limit_e { curve_recursion_limit = 32 };
inline double calc_sq_distance(double x1, double y1, double x2, double y2) {
  double dx = x2 - x1;
  double dy = y2 - y1;
  return dx * dx + dy * dy;
}

class curve3 {  // curve3 with its default approximation method curve_div (curve3_div)
 public:
  curve3() : m_approximation_scale(1.0), m_distance_tolerance_square(0.0), m_angle_tolerance(0.0), m_count(0) {}
  void reset() { m_points.clear(); m_count = 0; }
  void init(double x1, double y1, double x2, double y2, double x3, double y3) {
    m_points.clear();
    m_distance_tolerance_square = 0.5 / m_approximation_scale;
    m_distance_tolerance_square *= m_distance_tolerance_square;
    m_points.push_back(point_d(x1, y1));
    recursive_bezier(x1, y1, x2, y2, x3, y3, 0);
    m_points.push_back(point_d(x3, y3));
    m_count = 0;
  }
  void rewind(unsigned) { m_count = 0; }
  unsigned vertex(double* x, double* y) {
    if (m_count >= m_points.size()) return path_cmd_stop;
    const point_d& p = m_points[m_count++];
    *x = p.x;
    *y = p.y;
    return (m_count == 1) ? path_cmd_move_to : path_cmd_line_to;
  }
 private:
  void recursive_bezier(double x1, double y1, double x2, double y2, double x3, double y3, unsigned level) {
    if (level > curve_recursion_limit) return;
    double x12 = (x1 + x2) / 2;
    double y12 = (y1 + y2) / 2;
    double x23 = (x2 + x3) / 2;
    double y23 = (y2 + y3) / 2;
    double x123 = (x12 + x23) / 2;
    double y123 = (y12 + y23) / 2;
    double dx = x3 - x1;
    double dy = y3 - y1;
    double d = std::fabs(((x2 - x3) * dy - (y2 - y3) * dx));
    double da;
    if (d > curve_collinearity_epsilon) {
      if (d * d <= m_distance_tolerance_square * (dx * dx + dy * dy)) {
        if (m_angle_tolerance < curve_angle_tolerance_epsilon) {
          m_points.push_back(point_d(x123, y123));
          return;
        }
        da = std::fabs(std::atan2(y3 - y2, x3 - x2) - std::atan2(y2 - y1, x2 - x1));
        if (da >= pi) da = 2 * pi - da;
        if (da < m_angle_tolerance) {
          m_points.push_back(point_d(x123, y123));
          return;
        }
      }
    } else {
      da = dx * dx + dy * dy;
      if (da == 0) {
        d = calc_sq_distance(x1, y1, x2, y2);
      } else {
        d = ((x2 - x1) * dx + (y2 - y1) * dy) / da;
        if (d > 0 && d < 1) return;  // simple collinear case, 1---2---3
        if (d <= 0) d = calc_sq_distance(x2, y2, x1, y1);
        else if (d >= 1) d = calc_sq_distance(x2, y2, x3, y3);
        else d = calc_sq_distance(x2, y2, x1 + d * dx, y1 + d * dy);
      }
      if (d < m_distance_tolerance_square) {
        m_points.push_back(point_d(x2, y2));
        return;
      }
    }
    recursive_bezier(x1, y1, x12, y12, x123, y123, level + 1);
    recursive_bezier(x123, y123, x23, y23, x3, y3, level + 1);
  }
  double m_approximation_scale, m_distance_tolerance_square, m_angle_tolerance;
  unsigned m_count;
  std::vector<point_d> m_points;
};

template <class VertexSource>
class conv_curve {  // curve4 commands never occur in the reference (curve4To is commented out, DataGenerator.cpp:505-510)
 public:
  explicit conv_curve(VertexSource& source) : m_source(&source), m_last_x(0.0), m_last_y(0.0) {}
  void rewind(unsigned path_id) {
    m_source->rewind(path_id);
    m_last_x = 0.0;
    m_last_y = 0.0;
    m_curve3.reset();
  }
  unsigned vertex(double* x, double* y) {
    if (!is_stop(m_curve3.vertex(x, y))) {
      m_last_x = *x;
      m_last_y = *y;
      return path_cmd_line_to;
    }
    double end_x = 0, end_y = 0;
    unsigned cmd = m_source->vertex(x, y);
    switch (cmd) {
      case path_cmd_curve3:
        m_source->vertex(&end_x, &end_y);
        m_curve3.init(m_last_x, m_last_y, *x, *y, end_x, end_y);
        m_curve3.vertex(x, y);  // first call returns path_cmd_move_to
        m_curve3.vertex(x, y);  // this is the first vertex of the curve
        cmd = path_cmd_line_to;
        break;
      default:
        break;
    }
    m_last_x = *x;
    m_last_y = *y;
    return cmd;
  }
 private:
  VertexSource* m_source;
  double m_last_x, m_last_y;
  curve3 m_curve3;
};

}  // namespace agg

#endif
