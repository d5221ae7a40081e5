// See flatten.hpp. "DG.cpp" = /root/reference/src/caffe/DataGenerator.cpp.
// AGG 2.4 / CImg behaviour restated from SURVEY App. B (neither library is vendored).
#include "flatten.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <stdexcept>

#include "affine.hpp"

namespace ofdg {

namespace {

const double kPi = 3.14159265358979323846;  // agg::pi == cimg::PI

struct Pt {
  double x, y;
};

// agg::curve3_div::recursive_bezier with approximation_scale 1, angle_tolerance 0
// (SURVEY App. B.3): distance tolerance^2 = 0.25, collinearity eps 1e-30, depth limit 32.
void subdivide_quadratic(double x1, double y1, double x2, double y2, double x3, double y3,
                         unsigned level, std::vector<Pt>& pts) {
  if (level > 32) return;
  double x12 = (x1 + x2) / 2;
  double y12 = (y1 + y2) / 2;
  double x23 = (x2 + x3) / 2;
  double y23 = (y2 + y3) / 2;
  double x123 = (x12 + x23) / 2;
  double y123 = (y12 + y23) / 2;
  double dx = x3 - x1;
  double dy = y3 - y1;
  double d = std::fabs(((x2 - x3) * dy - (y2 - y3) * dx));
  const double tol2 = 0.25;
  if (d > 1e-30) {
    if (d * d <= tol2 * (dx * dx + dy * dy)) {
      pts.push_back(Pt{x123, y123});
      return;
    }
  } else {
    double da = dx * dx + dy * dy;
    auto sqd = [](double ax, double ay, double bx, double by) {
      double ex = bx - ax, ey = by - ay;
      return ex * ex + ey * ey;
    };
    if (da == 0) {
      d = sqd(x1, y1, x2, y2);
    } else {
      d = ((x2 - x1) * dx + (y2 - y1) * dy) / da;
      if (d > 0 && d < 1) return;  // 1---2---3: the two end points suffice
      if (d <= 0)
        d = sqd(x2, y2, x1, y1);
      else if (d >= 1)
        d = sqd(x2, y2, x3, y3);
      else
        d = sqd(x2, y2, x1 + d * dx, y1 + d * dy);
    }
    if (d < tol2) {
      pts.push_back(Pt{x2, y2});
      return;
    }
  }
  subdivide_quadratic(x1, y1, x12, y12, x123, y123, level + 1, pts);
  subdivide_quadratic(x123, y123, x23, y23, x3, y3, level + 1, pts);
}

inline FlatVertex to_fixed(double x, double y) {  // ras_conv_int::upscale
  return FlatVertex{iround(x * 256.0), iround(y * 256.0)};
}

Affine from6(const double m[6]) { return Affine{m[0], m[1], m[2], m[3], m[4], m[5]}; }

void bbox_of(const FlatVertex* v, int n, int32_t box[4]) {
  int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
  for (int i = 0; i < n; ++i) {
    x0 = std::min(x0, v[i].x); x1 = std::max(x1, v[i].x);
    y0 = std::min(y0, v[i].y); y1 = std::max(y1, v[i].y);
  }
  box[0] = x0 >> 8; box[1] = y0 >> 8; box[2] = x1 >> 8; box[3] = y1 >> 8;
}

// setMotion, DG.cpp:312-322
Affine motion_of(const ofdg_blueprint& b) {
  Affine m;
  m.then(Affine::rotation(b.rot));
  m.then(Affine::scaling(b.scale));
  m.then(Affine::translation(b.trans_x, b.trans_y));
  return m;
}

// setIntrinsicTransform, DG.cpp:302-310
Affine intrinsic_of(float alpha, float xs, float ys) {
  Affine m;
  m.then(Affine::rotation(alpha));
  m.then(Affine::translation(xs, ys));
  return m;
}

// cimg::mod(float, float)
float cimg_modf(float x, float m) {
  const double dx = (double)x, dm = (double)m;
  return (float)(dx - dm * std::floor(dx / dm));
}

void prepare_background(const ofdg_blueprint& b, const FlattenConfig& cfg, const Affine& tex_inv,
                        int spread, BgPrep& p) {
  const int W = cfg.W, H = cfg.H, tw = 2 * W, th = 2 * H;
  p.tex = (int32_t)((unsigned)b.tex_id % (unsigned)cfg.n_tex);
  const int w = cfg.tex_info[p.tex].w, h = cfg.tex_info[p.tex].h;
  p.shift_x = b.tex_shift_x;
  p.shift_y = b.tex_shift_y;
  // CImg get_rotate(angle, 1, 3): SURVEY App. B.5
  const float nangle = cimg_modf(b.tex_rot, 360.0f);
  p.rot_identity = (nangle == 0.f) ? 1 : 0;
  // CImg rotates by exact multiples of 90 degrees without interpolation (axis swaps / flips). The parameter streams cannot
  // reach them (tex_rot stays within +-pi, taken as degrees); blueprints supplied from outside that ask for one are refused,
  // as the oracle refuses them, instead of being rendered with the unrotated texture.
  if (!p.rot_identity && cimg_modf(nangle, 90.0f) == 0)
    throw std::runtime_error("background tex_rot of exactly 90 / 180 / 270 degrees (CImg's orthogonal rotations) is not implemented");
  if (p.rot_identity) {
    p.ca = 1.f; p.sa = 0.f;
    p.rw = w; p.rh = h;
  } else {
    const float rad = (float)(nangle * kPi / 180.0);
    p.ca = (float)std::cos(rad);
    p.sa = (float)std::sin(rad);
    const float ux = std::fabs((unsigned)(w - 1) * p.ca), uy = std::fabs((unsigned)(w - 1) * p.sa),
                vx = std::fabs((unsigned)(h - 1) * p.sa), vy = std::fabs((unsigned)(h - 1) * p.ca);
    p.rw = (int)std::floor((1 + ux + vx) + 0.5f);
    p.rh = (int)std::floor((1 + uy + vy) + 0.5f);
  }
  p.w2 = 0.5f * (unsigned)(w - 1);
  p.h2 = 0.5f * (unsigned)(h - 1);
  p.rw2 = 0.5f * (unsigned)(p.rw - 1);
  p.rh2 = 0.5f * (unsigned)(p.rh - 1);
  if (w >= tw && h >= th) {
    // crop(width/2-tex_w/2, height/2-tex_h/2, width/2-tex_w/2+tex_w/zoom-1, ..., 3), DG.cpp:99-102
    const float zoom = b.tex_scale;
    const int x0 = w / 2 - tw / 2, y0 = h / 2 - th / 2;
    const int x1 = (int)(w / 2 - tw / 2 + tw / zoom - 1);
    const int y1 = (int)(h / 2 - th / 2 + th / zoom - 1);
    p.crop_x0 = std::min(x0, x1);
    p.crop_y0 = std::min(y0, y1);
    p.crop_w = std::abs(x1 - x0) + 1;
    p.crop_h = std::abs(y1 - y0) + 1;
  } else {
    // a texture smaller than 2W x 2H is not cropped: the whole rotated image is resized (DG.cpp:103-107)
    p.crop_x0 = 0; p.crop_y0 = 0;
    p.crop_w = p.rw; p.crop_h = p.rh;
  }
  if (p.crop_w < 2 || p.crop_h < 2) throw std::runtime_error("background texture crop degenerates to less than 2 pixels");
  if (p.crop_w > 40 * tw || p.crop_h > 40 * th) throw std::runtime_error("background texture more than 40x larger than the prepared size");
  p.general = (p.crop_w * 10 > tw * 13 || p.crop_h * 10 > th * 13) ? 1 : 0;
  p.pad = 0;

  // Part of the prepared texture the renderer touches: the centre W x H window (frame 0)
  // plus the footprint of the frame-1 warp (4 taps around tex_inv * pixel centre). A background with a warp field samples the
  // warped canvas up to `spread` pixels (twice the inverse field's reach, + 2 for the taps) outside the window -- never
  // outside the canvas (Dirichlet boundary); spread < 0: reach unknown, the whole canvas.
  int nx0 = W / 2, ny0 = H / 2, nx1 = W / 2 + W - 1, ny1 = H / 2 + H - 1;
  if (spread < 0) {
    nx0 = 0; ny0 = 0; nx1 = tw - 1; ny1 = th - 1;
  } else {
    const double cx[2] = {std::max(0.0, W / 2 + 0.0 - spread), std::min((double)tw, W / 2 + W + 1.0 + spread)};
    const double cy[2] = {std::max(0.0, H / 2 + 0.0 - spread), std::min((double)th, H / 2 + H + 1.0 + spread)};
    double fx0 = 1e300, fy0 = 1e300, fx1 = -1e300, fy1 = -1e300;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        double x = cx[i], y = cy[j];
        tex_inv.apply(&x, &y);
        fx0 = std::min(fx0, x); fx1 = std::max(fx1, x);
        fy0 = std::min(fy0, y); fy1 = std::max(fy1, y);
      }
    const int ax0 = (int)std::floor(fx0) - 3, ax1 = (int)std::ceil(fx1) + 3;
    const int ay0 = (int)std::floor(fy0) - 3, ay1 = (int)std::ceil(fy1) + 3;
    if (ax0 < 0 || ax1 > tw - 1) { nx0 = 0; nx1 = tw - 1; } else { nx0 = std::min(nx0, ax0); nx1 = std::max(nx1, ax1); }
    if (ay0 < 0 || ay1 > th - 1) { ny0 = 0; ny1 = th - 1; } else { ny0 = std::min(ny0, ay0); ny1 = std::max(ny1, ay1); }
  }
  p.need[0] = nx0; p.need[1] = ny0; p.need[2] = nx1; p.need[3] = ny1;
}

// One outline of blueprint `b` (an ellipse or a polygon), both frames.
void realize_shape(const ofdg_task_batch& tb, const ofdg_blueprint& b, const Affine& bg_n,
                   FlatBatch& out, Affine* motion_out, int field, int reach) {
  Affine I = intrinsic_of(b.init_rot, b.init_trans_x, b.init_trans_y);
  Affine M = motion_of(b);
  M.then(bg_n);  // addBackgroundMotion, DG.cpp:324-335
  Affine IM = I;
  IM.then(M);    // renderMasks: save = intrinsic; save *= motion (DG.cpp:469-470, 522-523)
  if (motion_out) *motion_out = M;

  FlatShape s{};
  s.additive = b.is_additive_component ? 1 : 0;
  const Affine* tf[2] = {&I, &IM};
  for (int f = 0; f < 2; ++f) {
    double m[6];
    tf[f]->store(m);
    s.vbegin[f] = (int32_t)out.verts.size();
    if (b.obj_type == OFDG_OBJ_ELLIPSE) {
      flatten_ellipse(b.ellipse_scale_x, b.ellipse_scale_y, m, out.verts);
    } else if (b.obj_type == OFDG_OBJ_POLYGON) {
      if (b.seg_count < 1 || b.seg_begin < 0 || b.seg_begin + b.seg_count > tb.n_segments)
        throw std::runtime_error("polygon blueprint with a bad segment range");
      flatten_polygon(tb.seg_type + b.seg_begin, tb.seg_x + b.seg_begin, tb.seg_y + b.seg_begin,
                      b.seg_count, m, out.verts);
    } else {
      throw std::runtime_error("(RealizeObjectBlueprint) Bad object type, or not intended in this mode");  // DG.cpp:1143
    }
    s.vcount[f] = (int32_t)out.verts.size() - s.vbegin[f];
    bbox_of(out.verts.data() + s.vbegin[f], s.vcount[f], s.bbox[f]);
  }
  s.deform = -1;
  for (int i = 0; i < 4; ++i) s.raw1[i] = s.bbox[1][i];
  if (field >= 0) {
    // MovingObjectBase::renderMasks warps this outline's frame-1 masks by the inverse field (DG.cpp:370-386):
    // out(x,y) = in((x,y) + iflow(x,y)) can be non-zero up to `reach` pixels away from the outline.
    s.deform = (int32_t)out.deform_shape.size();
    out.deform_shape.push_back((int32_t)out.shapes.size());
    out.deform_field.push_back(field);
    s.bbox[1][0] -= reach + 2; s.bbox[1][1] -= reach + 2; s.bbox[1][2] += reach + 2; s.bbox[1][3] += reach + 2;
  }
  out.shapes.push_back(s);
}

}  // namespace

// agg::ellipse::vertex through agg::conv_transform: 100 steps, ccw.
void flatten_ellipse(double rx, double ry, const double m6[6], std::vector<FlatVertex>& out) {
  const Affine m = from6(m6);
  const unsigned num = 100;  // setEllipse(0, 0, rx, ry, 100), DG.cpp:1080
  // the 100 step angles do not depend on the ellipse: cos/sin evaluated once (same libm, same values)
  static const struct Circle {
    double c[100], s[100];
    Circle() {
      for (unsigned step = 0; step < 100; ++step) {
        double angle = double(step) / double(100) * 2.0 * kPi;
        c[step] = std::cos(angle);
        s[step] = std::sin(angle);
      }
    }
  } circle;
  out.reserve(out.size() + num);
  for (unsigned step = 0; step < num; ++step) {
    double x = 0.0 + circle.c[step] * rx;
    double y = 0.0 + circle.s[step] * ry;
    m.apply(&x, &y);
    out.push_back(to_fixed(x, y));
  }
}

// path_storage -> conv_transform -> conv_curve (DG.cpp:491-531, 1091-1114): the path's
// vertices (curve control points included) are transformed first, curves are then
// subdivided in screen space.
void flatten_polygon(const int32_t* seg_type, const float* seg_x, const float* seg_y, int n,
                     const double m6[6], std::vector<FlatVertex>& out) {
  const Affine m = from6(m6);
  auto tp = [&](int i) {
    Pt p{(double)seg_x[i], (double)seg_y[i]};
    m.apply(&p.x, &p.y);
    return p;
  };
  Pt last = tp(0);  // resetPath: move_to
  out.push_back(to_fixed(last.x, last.y));
  std::vector<Pt> pts;
  for (int i = 1; i < n; ++i) {
    switch (seg_type[i]) {
      case OFDG_SEG_LINE: {
        last = tp(i);
        out.push_back(to_fixed(last.x, last.y));
        break;
      }
      case OFDG_SEG_CURVE3: {
        if (i + 1 >= n) throw std::runtime_error("curve3 segment without an end point");
        Pt c = tp(i), e = tp(i + 1);
        pts.clear();
        subdivide_quadratic(last.x, last.y, c.x, c.y, e.x, e.y, 0, pts);
        pts.push_back(e);  // curve3_div::bezier appends the end point; the start is swallowed by conv_curve
        for (const Pt& p : pts) out.push_back(to_fixed(p.x, p.y));
        last = e;
        ++i;
        break;
      }
      default:
        throw std::runtime_error("PolySegmentType_t::Dummy found, this should have been skipped!");  // DG.cpp:1096
    }
  }
}

void flatten(const ofdg_task_batch& tb, const FlattenConfig& cfg, FlatBatch& out) {
  const int W = cfg.W, H = cfg.H;
  if (cfg.n_tex <= 0) throw std::runtime_error("texture pool is empty");
  for (int t = 0; t < tb.n_tasks; ++t) {
    const int b0 = tb.task_begin[t], b1 = tb.task_begin[t + 1];
    if (b1 <= b0) throw std::runtime_error("task without a background blueprint");
    const ofdg_blueprint& bg = tb.blueprints[b0];

    FlatSample smp{};
    // Background: I = T(W, H) (DG.cpp:662), texture warped by I^-1 * M * I on the 2W x 2H canvas.
    const Affine bgI = intrinsic_of(0.f, (float)W, (float)H);
    const Affine bgM = motion_of(bg);
    Affine tex_tf = bgI.inverse();
    tex_tf.then(bgM);
    tex_tf.then(bgI);
    const Affine tex_inv = tex_tf.inverse();
    tex_inv.store(smp.bg_tex_inv);
    bgM.store(smp.bg_motion);
    bgM.inverse().store(smp.bg_motion_inv);
    const bool bg_deformed = (cfg.mode == 9 && bg.do_warpfield_deformation && bg.field_id >= 0);
    smp.bg_field = bg_deformed ? bg.field_id : -1;
    if (smp.bg_field >= cfg.n_fields) throw std::runtime_error("background refers to a warp field that was not injected (ofdg_set_fields)");
    // (a reach beyond the canvas is the whole canvas anyway: clamped so that 2 * reach + 2 cannot overflow)
    const int spread = !bg_deformed ? 0 : (cfg.field_reach ? 2 * std::min(cfg.field_reach[bg.field_id], 4 * (W + H)) + 2 : -1);
    prepare_background(bg, cfg, tex_inv, spread, smp.prep);
    if (tb.augment) smp.aug = tb.augment[t];

    // addBackgroundMotion's bracket T(-W/2,-H/2) * M_bg * T(W/2,H/2), DG.cpp:327-329
    Affine bg_n = Affine::translation(-W / 2., -H / 2.);
    bg_n.then(bgM);
    bg_n.then(Affine::translation(W / 2., H / 2.));

    smp.obj_begin = (int32_t)out.objects.size();
    for (int bi = b0 + 1; bi < b1; ++bi) {
      const ofdg_blueprint& b = tb.blueprints[bi];
      if (b.parent >= 0) continue;  // components are realised with their composite
      FlatObject o{};
      o.obj_id = b.obj_id;
      o.tex = (int32_t)((unsigned)b.tex_id % (unsigned)cfg.n_tex);
      o.field = (cfg.mode == 9 && b.do_warpfield_deformation && b.field_id >= 0) ? b.field_id : -1;
      if (o.field >= cfg.n_fields) throw std::runtime_error("blueprint refers to a warp field that was not injected (ofdg_set_fields)");
      const int reach = (o.field >= 0 && cfg.field_reach) ? cfg.field_reach[o.field] : 0;
      o.shape_begin = (int32_t)out.shapes.size();
      Affine M;
      if (b.obj_type == OFDG_OBJ_COMPOSITE) {
        o.composite = 1;
        if (b.comp_begin < b0 || b.comp_begin + b.comp_count > b1 || b.comp_count < 0)
          throw std::runtime_error("composite blueprint with a bad component range");
        for (int ci = 0; ci < b.comp_count; ++ci) {
          const ofdg_blueprint& c = tb.blueprints[b.comp_begin + ci];
          // components carry the parent's field (DG.cpp:1157-1163)
          realize_shape(tb, c, bg_n, out, nullptr, o.field, reach);
        }
        // the composite's own motion drives its texture and flow (DG.cpp:1151-1155)
        M = motion_of(b);
        M.then(bg_n);
      } else {
        realize_shape(tb, b, bg_n, out, &M, o.field, reach);
      }
      o.shape_count = (int32_t)out.shapes.size() - o.shape_begin;
      M.store(o.motion);
      M.inverse().store(o.tex_inv);
      for (int f = 0; f < 2; ++f) {
        int32_t* bb = o.bbox[f];
        bb[0] = INT_MAX; bb[1] = INT_MAX; bb[2] = INT_MIN; bb[3] = INT_MIN;
        for (int si = 0; si < o.shape_count; ++si) {
          const int32_t* sb = out.shapes[o.shape_begin + si].bbox[f];
          bb[0] = std::min(bb[0], sb[0]); bb[1] = std::min(bb[1], sb[1]);
          bb[2] = std::max(bb[2], sb[2]); bb[3] = std::max(bb[3], sb[3]);
        }
      }
      out.objects.push_back(o);
    }
    smp.obj_count = (int32_t)out.objects.size() - smp.obj_begin;
    if (smp.obj_count > 254) throw std::runtime_error("more than 254 foreground objects in one sample (object ids are tracked in one byte per pixel)");
    out.samples.push_back(smp);
  }
}

}  // namespace ofdg
