// CPU ORACLE -- TEST INFRASTRUCTURE (see oracle.h).
//
// Producer side of the reference's mode-9 non-rigid warp fields, restated from
//   /root/reference/src/caffe/WarpFields.cpp  ("WF.cpp"):
//   Supports::Gaussian2D (WF.cpp:88-112), Displacers::{Translation,Rotation,Zoom} (WF.cpp:191-260),
//   DisplacementComposer (WF.cpp:296-316), FlowField::init_from_DisplacementComposer -- 17 ping-pong
//   self-compositions forward and inverse with out-of-bounds flagging -> NaN (WF.cpp:337-437),
//   clamp_near_zeros (WF.cpp:444-455) and CropGenerator::worker_thread_loop's 9x7 hex grid of random
//   displacers on a 3*max(W,H) canvas with 8x5 crops of (W+1)x(H+1) (WF.cpp:540-641).
// The reference seeds this from std::random_device (not reproducible by construction), so the fields
// are INPUTS to every parity test; here the engine is seeded explicitly. The order in which the
// reference's constructor arguments draw from the engine is unspecified in C++; left-to-right is used.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "oracle.h"

namespace wf {

struct FImg {  // CImg<float> with 2 channels
  int w = 0, h = 0;
  std::vector<float> d;
  FImg() {}
  FImg(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_ * 2, 0.f) {}
  float& at(int x, int y, int c) { return d[(size_t)x + (size_t)y * w + (size_t)c * w * h]; }
  float at(int x, int y, int c) const { return d[(size_t)x + (size_t)y * w + (size_t)c * w * h]; }
  // CImg::_linear_atXY (Neumann)
  float linear(float fx, float fy, int c) const {
    const float nfx = fx <= 0 ? 0 : (fx >= w - 1 ? (float)(w - 1) : fx), nfy = fy <= 0 ? 0 : (fy >= h - 1 ? (float)(h - 1) : fy);
    const unsigned int x = (unsigned int)nfx, y = (unsigned int)nfy;
    const float dx = nfx - x, dy = nfy - y;
    const unsigned int nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
    const float Icc = at(x, y, c), Inc = at(nx, y, c), Icn = at(x, ny, c), Inn = at(nx, ny, c);
    return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
  }
};

struct Gaussian2D {  // WF.cpp:88-112
  float cx, cy, a, b, c, d, ratio_x_y, sigma_sq, gauss_prefactor, normalizer;
  Gaussian2D(float cx_, float cy_, float sigma_x, float sigma_y, float angle)
      : cx(cx_), cy(cy_), a(std::cos(angle)), b(-std::sin(angle)), c(std::sin(angle)), d(std::cos(angle)),
        ratio_x_y(sigma_x / sigma_y), sigma_sq(sigma_x * sigma_x), gauss_prefactor(1 / std::sqrt(2 * M_PI * sigma_sq)), normalizer(1) {
    normalizer = 1 / raw_at(cx, cy);
  }
  float raw_at(float x, float y) const {
    const float rx = a * (x - cx) + b * (y - cy);
    const float ry = (c * (x - cx) + d * (y - cy)) * ratio_x_y;
    const float dist_sq{rx * rx + ry * ry};
    return gauss_prefactor * std::exp(-dist_sq / (2 * sigma_sq));
  }
  float at(float x, float y) const { return normalizer * raw_at(x, y); }
};

struct Displacer {  // WF.cpp:191-260
  int kind;  // 0 translation, 1 rotation, 2 zoom
  float cx = 0, cy = 0, dx = 0, dy = 0, sin_o = 0, cos_o = 0, sin_no = 0, cos_no = 0, factor = 1, ifactor = 1;
  std::unique_ptr<Gaussian2D> support;
  void raw_flow(float x, float y, float& fx, float& fy) const {
    if (kind == 0) { fx = dx; fy = dy; return; }
    const float ddx{x - cx}, ddy{y - cy};
    if (kind == 1) {
      const float rot_dx{cos_no * ddx - sin_no * ddy}, rot_dy{sin_no * ddx + cos_no * ddy};
      fx = rot_dx - ddx; fy = rot_dy - ddy;
    } else { fx = factor * ddx - ddx; fy = factor * ddy - ddy; }
  }
  void raw_iflow(float x, float y, float& fx, float& fy) const {
    if (kind == 0) { fx = -dx; fy = -dy; return; }
    const float ddx{x - cx}, ddy{y - cy};
    if (kind == 1) {
      const float rot_dx{cos_o * ddx - sin_o * ddy}, rot_dy{sin_o * ddx + cos_o * ddy};
      fx = rot_dx - ddx; fy = rot_dy - ddy;
    } else { fx = ifactor * ddx - ddx; fy = ifactor * ddy - ddy; }
  }
};

// FlowField::init_from_DisplacementComposer's doubling loop (WF.cpp:366-398 / 406-434)
static void compose(FImg& field, int W, int H) {
  FImg tmp = field;
  std::vector<unsigned char> flagged((size_t)W * H, 0);
  for (int iter = 17; iter > 0; --iter) {
    FImg& from = (iter % 2 == 1 ? tmp : field);
    FImg& to = (iter % 2 == 1 ? field : tmp);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const float fx = from.at(x, y, 0), fy = from.at(x, y, 1);
        if (x + fx < 0 || x + fx >= W || y + fy < 0 || y + fy >= H) {
          flagged[(size_t)y * W + x] = 255;
          to.at(x, y, 0) = fx; to.at(x, y, 1) = fy;
          continue;
        }
        to.at(x, y, 0) = fx + from.linear(x + fx, y + fy, 0);
        to.at(x, y, 1) = fy + from.linear(x + fx, y + fy, 1);
      }
  }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      if (x + field.at(x, y, 0) < 0 || x + field.at(x, y, 0) >= W || y + field.at(x, y, 1) < 0 || y + field.at(x, y, 1) >= H)
        flagged[(size_t)y * W + x] = 255;
      if (flagged[(size_t)y * W + x]) {
        field.at(x, y, 0) = std::numeric_limits<float>::quiet_NaN();
        field.at(x, y, 1) = std::numeric_limits<float>::quiet_NaN();
      }
    }
}

}  // namespace wf

extern "C" int oracle_generate_fields(int32_t W, int32_t H, uint32_t seed, int32_t n_fields, float* out) {
  try {
    using namespace wf;
    std::mt19937 mersenne(seed);
    std::uniform_int_distribution<> displacer_type(0, 2);
    std::uniform_real_distribution<> generic_param(-1, 1);
    const int big_size{std::max(W, H) * 3};
    const size_t per_field = (size_t)2 * 2 * (H + 1) * (W + 1);
    int produced = 0;
    while (produced < n_fields) {
      std::vector<Displacer> ds;
      const int spacing{200};
      const int isosceles_spacing{(int)(spacing / 2. * std::sqrt(3.))};
      const int rows{(big_size + isosceles_spacing - 1) / isosceles_spacing};
      const int cols{big_size / spacing};
      for (int yidx = 0; yidx < rows; ++yidx)
        for (int xidx = 0; xidx < cols; ++xidx) {
          const int x = xidx * spacing + (yidx % 2 == 1 ? spacing / 2 : 0) + spacing / 2;
          const int y = yidx * isosceles_spacing + spacing / 2;
          Displacer d;
          d.kind = displacer_type(mersenne);
          auto g = [&]() { return generic_param(mersenne); };
          if (d.kind == 0) { d.dx = (float)(g() * 3e-4); d.dy = (float)(g() * 3e-4); }
          else if (d.kind == 1) {
            d.cx = (float)(x + g() * 10); d.cy = (float)(y + g() * 10);
            const float omega = (float)(g() * M_PI * 2e-6);
            d.sin_o = std::sin(omega); d.cos_o = std::cos(omega); d.sin_no = std::sin(-omega); d.cos_no = std::cos(-omega);
          } else {
            d.cx = (float)(x + g() * 10); d.cy = (float)(y + g() * 10);
            d.factor = (float)(1 + g() * 2e-6); d.ifactor = (float)(1. / d.factor);
          }
          const float scx = (float)(x + g() * 10), scy = (float)(y + g() * 10), ssx = (float)(50 + g() * 20), ssy = (float)(50 + g() * 20),
                      sang = (float)(g() * M_PI);
          d.support.reset(new Gaussian2D(scx, scy, ssx, ssy, sang));
          ds.push_back(std::move(d));
        }
      FImg flow(big_size, big_size), iflow(big_size, big_size);
      for (int y = 0; y < big_size; ++y)  // WF.cpp:347-354
        for (int x = 0; x < big_size; ++x) {
          float fx = 0, fy = 0, ix = 0, iy = 0;
          for (const Displacer& d : ds) {
            const float w{d.support->at((float)x, (float)y)};
            float a, b;
            d.raw_flow((float)x, (float)y, a, b);
            fx += a * w; fy += b * w;
            d.raw_iflow((float)x, (float)y, a, b);
            ix += a * w; iy += b * w;
          }
          flow.at(x, y, 0) = fx; flow.at(x, y, 1) = fy;
          iflow.at(x, y, 0) = ix; iflow.at(x, y, 1) = iy;
        }
      compose(flow, big_size, big_size);
      compose(iflow, big_size, big_size);
      const float threshold{1e-3};  // clamp_near_zeros, WF.cpp:444-455 (NaN compares false and stays)
      for (float& v : flow.d) if (std::abs(v) < threshold) v = 0.f;
      for (float& v : iflow.d) if (std::abs(v) < threshold) v = 0.f;
      // crops: get_crop(x, y, x+W, y+H) is inclusive -> (W+1) x (H+1)   (WF.cpp:619-634)
      for (int y = H / 4; y < big_size - 5 * H / 4 && produced < n_fields; y += H / 3)
        for (int x = W / 4; x < big_size - 5 * W / 4 && produced < n_fields; x += W / 3) {
          float* dst = out + (size_t)produced * per_field;
          const FImg* src[2] = {&flow, &iflow};
          for (int k = 0; k < 2; ++k)
            for (int c = 0; c < 2; ++c)
              for (int yy = 0; yy <= H; ++yy)
                for (int xx = 0; xx <= W; ++xx) *dst++ = src[k]->at(x + xx, y + yy, c);
          ++produced;
        }
    }
    return 0;
  } catch (const std::exception& e) {
    return 1;
  }
}
