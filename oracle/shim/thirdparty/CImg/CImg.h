// TEST INFRASTRUCTURE, NOT PRODUCT CODE (oracle/shim).
//
// A from-scratch stand-in for the part of CImg (>= 2.0.0, http://cimg.eu) the reference generator calls.
// The reference does not vendor CImg ("download CImg.h from cimg.eu",
// /root/reference/include/thirdparty/download-__CImg.h__-from-__cimg.eu__.txt) and it is not available
// offline, so this header RESTATES the published CImg algorithms behind CImg's own names and signatures,
// just far enough for /root/reference/src/caffe/DataGenerator.cpp and WarpFields.cpp to compile untouched
// (call sites: DataGenerator.cpp:97-107, 128-131, 180, 228, 245, 374-383, 404-405, 680-681, 715-716, 754-759,
// 780-792, 1197-1200, 1229-1245; WarpFields.cpp:341-342, 347-455, 623-624). Semantics follow SURVEY.md
// App. B.5. PARITY OF THIS PART REMAINS UNPINNED (CImg's boundary-3 "mirror" handling of shift / rotate /
// crop is version dependent); put a real CImg.h first on the include path (oracle/ref_build.sh
// CIMG_INCLUDE=...) to replace it.
#ifndef OFDG_ORACLE_CIMG_SHIM_H_
#define OFDG_ORACLE_CIMG_SHIM_H_

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>

#define cimg_version 299  /* "a CImg 2.x": the shim has no version of its own */

#define cimg_forX(img, x) for (int x = 0; x < (int)((img)._width); ++x)
#define cimg_forY(img, y) for (int y = 0; y < (int)((img)._height); ++y)
#define cimg_forZ(img, z) for (int z = 0; z < (int)((img)._depth); ++z)
#define cimg_forC(img, c) for (int c = 0; c < (int)((img)._spectrum); ++c)
#define cimg_forXY(img, x, y) cimg_forY(img, y) cimg_forX(img, x)
#define cimg_forXYC(img, x, y, c) cimg_forC(img, c) cimg_forXY(img, x, y)
#define cimg_forXYZC(img, x, y, z, c) cimg_forC(img, c) cimg_forZ(img, z) cimg_forXY(img, x, y)
#define cimg_forYZC(img, y, z, c) cimg_forC(img, c) cimg_forZ(img, z) cimg_forY(img, y)
#define cimg_forXZC(img, x, z, c) cimg_forC(img, c) cimg_forZ(img, z) cimg_forX(img, x)

namespace cimg_library {

struct CImgException : public std::runtime_error { explicit CImgException(const std::string& m) : std::runtime_error(m) {} };
struct CImgArgumentException : public CImgException { explicit CImgArgumentException(const std::string& m) : CImgException(m) {} };
struct CImgIOException : public CImgException { explicit CImgIOException(const std::string& m) : CImgException(m) {} };

namespace cimg {
const double PI = 3.14159265358979323846;
inline int mod(const int x, const int m) { return x >= 0 ? x % m : (x % m ? m + x % m : 0); }
inline float mod(const float x, const float m) {
  const double dx = (double)x, dm = (double)m;
  return (float)(dx - dm * std::floor(dx / dm));
}
template <class T> inline T abs(const T& a) { return a >= 0 ? a : -a; }
inline float round(const float x) { return (float)std::floor(x + 0.5f); }
template <class T> inline T cut(const T& val, const T& val_min, const T& val_max) { return val < val_min ? val_min : val > val_max ? val_max : val; }
template <class T> struct superset_float { typedef float type; };
template <> struct superset_float<double> { typedef double type; };
}  // namespace cimg

template <class T>
struct CImg {
  typedef typename cimg::superset_float<T>::type Tfloat;
  unsigned int _width, _height, _depth, _spectrum;
  bool _is_shared;
  T* _data;

  // ---- construction / assignment
  CImg() : _width(0), _height(0), _depth(0), _spectrum(0), _is_shared(false), _data(0) {}
  explicit CImg(const unsigned int w, const unsigned int h = 1, const unsigned int d = 1, const unsigned int s = 1)
      : _width(0), _height(0), _depth(0), _spectrum(0), _is_shared(false), _data(0) { assign(w, h, d, s); }
  CImg(const unsigned int w, const unsigned int h, const unsigned int d, const unsigned int s, const T& value)
      : _width(0), _height(0), _depth(0), _spectrum(0), _is_shared(false), _data(0) { assign(w, h, d, s); fill(value); }
  CImg(const T* values, const unsigned int w, const unsigned int h = 1, const unsigned int d = 1, const unsigned int s = 1, const bool is_shared = false)
      : _width(0), _height(0), _depth(0), _spectrum(0), _is_shared(false), _data(0) {
    if (is_shared) { _width = w; _height = h; _depth = d; _spectrum = s; _is_shared = true; _data = const_cast<T*>(values); }
    else { assign(w, h, d, s); if (size()) std::memcpy(_data, values, size() * sizeof(T)); }
  }
  CImg(const CImg<T>& img) : _width(0), _height(0), _depth(0), _spectrum(0), _is_shared(false), _data(0) {
    if (img._is_shared) { _width = img._width; _height = img._height; _depth = img._depth; _spectrum = img._spectrum; _is_shared = true; _data = img._data; }
    else { assign(img._width, img._height, img._depth, img._spectrum); if (size()) std::memcpy(_data, img._data, size() * sizeof(T)); }
  }
  CImg(CImg<T>&& img) : _width(img._width), _height(img._height), _depth(img._depth), _spectrum(img._spectrum), _is_shared(img._is_shared), _data(img._data) {
    img._width = img._height = img._depth = img._spectrum = 0; img._is_shared = false; img._data = 0;
  }
  ~CImg() { if (!_is_shared) delete[] _data; }
  CImg<T>& operator=(const CImg<T>& img) {
    if (this == &img) return *this;
    if (_is_shared) { _is_shared = false; _data = 0; _width = _height = _depth = _spectrum = 0; }
    assign(img._width, img._height, img._depth, img._spectrum);
    if (size()) std::memcpy(_data, img._data, size() * sizeof(T));
    return *this;
  }
  CImg<T>& operator=(CImg<T>&& img) { swap(img); return *this; }
  CImg<T>& swap(CImg<T>& img) {
    std::swap(_width, img._width); std::swap(_height, img._height); std::swap(_depth, img._depth); std::swap(_spectrum, img._spectrum);
    std::swap(_is_shared, img._is_shared); std::swap(_data, img._data);
    return img;
  }
  CImg<T>& move_to(CImg<T>& img) { img.swap(*this); assign(); return img; }
  CImg<T>& assign() {
    if (!_is_shared) delete[] _data;
    _width = _height = _depth = _spectrum = 0; _is_shared = false; _data = 0;
    return *this;
  }
  CImg<T>& assign(const unsigned int w, const unsigned int h = 1, const unsigned int d = 1, const unsigned int s = 1) {
    const size_t siz = (size_t)w * h * d * s;
    if (!siz) return assign();
    if (siz != size() || _is_shared) {
      if (!_is_shared) delete[] _data;
      _is_shared = false;
      _data = new T[siz];
    }
    _width = w; _height = h; _depth = d; _spectrum = s;
    return *this;
  }

  // ---- accessors
  int width() const { return (int)_width; }
  int height() const { return (int)_height; }
  int depth() const { return (int)_depth; }
  int spectrum() const { return (int)_spectrum; }
  size_t size() const { return (size_t)_width * _height * _depth * _spectrum; }
  bool is_empty() const { return !(_data && _width && _height && _depth && _spectrum); }
  operator bool() const { return !is_empty(); }
  T* data() { return _data; }
  const T* data() const { return _data; }
  T* data(const unsigned int x, const unsigned int y = 0, const unsigned int z = 0, const unsigned int c = 0) { return _data + offset(x, y, z, c); }
  const T* data(const unsigned int x, const unsigned int y = 0, const unsigned int z = 0, const unsigned int c = 0) const { return _data + offset(x, y, z, c); }
  size_t offset(const unsigned int x, const unsigned int y, const unsigned int z, const unsigned int c) const {
    return x + (size_t)y * _width + (size_t)z * _width * _height + (size_t)c * _width * _height * _depth;
  }
  // No bounds checks, exactly like CImg: the reference indexes channels through the z slot of depth-1
  // images (DataGenerator.cpp:130, 245, 404-405), which lands on the channel plane.
  T& operator()(const unsigned int x, const unsigned int y = 0, const unsigned int z = 0, const unsigned int c = 0) { return _data[offset(x, y, z, c)]; }
  const T& operator()(const unsigned int x, const unsigned int y = 0, const unsigned int z = 0, const unsigned int c = 0) const { return _data[offset(x, y, z, c)]; }
  T atXY(const int x, const int y, const int z, const int c, const T& out_value) const {
    return (x < 0 || y < 0 || x >= width() || y >= height()) ? out_value : (*this)(x, y, z, c);
  }

  CImg<T>& fill(const T& val) { for (size_t i = 0, n = size(); i < n; ++i) _data[i] = val; return *this; }
  template <class t> CImg<T>& operator*=(const t value) { for (size_t i = 0, n = size(); i < n; ++i) _data[i] = (T)(_data[i] * value); return *this; }

  // ---- interpolated reads
  // _linear_atXY / linear_atXY(fx, fy, z, c): bilinear with clamped (Neumann) coordinates
  Tfloat _linear_atXY(const float fx, const float fy, const int z = 0, const int c = 0) const {
    const float nfx = cimg::cut(fx, 0.f, (float)(width() - 1)), nfy = cimg::cut(fy, 0.f, (float)(height() - 1));
    const unsigned int x = (unsigned int)nfx, y = (unsigned int)nfy;
    const float dx = nfx - x, dy = nfy - y;
    const unsigned int nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
    const Tfloat Icc = (Tfloat)(*this)(x, y, z, c), Inc = (Tfloat)(*this)(nx, y, z, c), Icn = (Tfloat)(*this)(x, ny, z, c), Inn = (Tfloat)(*this)(nx, ny, z, c);
    return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
  }
  Tfloat linear_atXY(const float fx, const float fy, const int z = 0, const int c = 0) const {
    if (is_empty()) throw CImgArgumentException("linear_atXY(): empty instance");
    return _linear_atXY(fx, fy, z, c);
  }
  // linear_atXY(fx, fy, z, c, out_value): bilinear, out_value outside (Dirichlet). NaN / huge coordinates: the
  // x86-64 build of the real thing converts them to INT_MIN, so every tap is out of range; spelled out here
  // because the conversion is undefined behaviour in C++.
  Tfloat linear_atXY(const float fx, const float fy, const int z, const int c, const T& out_value) const {
    const int x = to_int(fx) - (fx >= 0 ? 0 : 1), nx = x + 1, y = to_int(fy) - (fy >= 0 ? 0 : 1), ny = y + 1;
    const float dx = fx - x, dy = fy - y;
    const Tfloat Icc = (Tfloat)atXY(x, y, z, c, out_value), Inc = (Tfloat)atXY(nx, y, z, c, out_value),
                 Icn = (Tfloat)atXY(x, ny, z, c, out_value), Inn = (Tfloat)atXY(nx, ny, z, c, out_value);
    return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
  }

  // ---- geometry
  CImg<T> get_crop(const int x0, const int y0, const int x1, const int y1, const unsigned int boundary_conditions = 0) const {
    if (is_empty()) throw CImgArgumentException("crop(): empty instance");
    const int nx0 = x0 < x1 ? x0 : x1, nx1 = x0 ^ x1 ^ nx0, ny0 = y0 < y1 ? y0 : y1, ny1 = y0 ^ y1 ^ ny0;
    CImg<T> res(1U + nx1 - nx0, 1U + ny1 - ny0, _depth, _spectrum);
    const int w2 = 2 * width(), h2 = 2 * height();
    cimg_forXYZC(res, x, y, z, c) {
      const int sx = nx0 + x, sy = ny0 + y;
      if (sx >= 0 && sy >= 0 && sx < width() && sy < height()) { res(x, y, z, c) = (*this)(sx, sy, z, c); continue; }
      switch (boundary_conditions) {
        case 3: {  // mirror
          const int mx = cimg::mod(sx, w2), my = cimg::mod(sy, h2);
          res(x, y, z, c) = (*this)(mx < width() ? mx : w2 - mx - 1, my < height() ? my : h2 - my - 1, z, c);
        } break;
        case 2: res(x, y, z, c) = (*this)(cimg::mod(sx, width()), cimg::mod(sy, height()), z, c); break;
        case 1: res(x, y, z, c) = (*this)(sx < 0 ? 0 : (sx >= width() ? width() - 1 : sx), sy < 0 ? 0 : (sy >= height() ? height() - 1 : sy), z, c); break;
        default: res(x, y, z, c) = (T)0;
      }
    }
    return res;
  }
  CImg<T>& crop(const int x0, const int y0, const int x1, const int y1, const unsigned int boundary_conditions = 0) {
    return get_crop(x0, y0, x1, y1, boundary_conditions).move_to(*this);
  }

  CImg<T> get_shift(const int delta_x, const int delta_y = 0, const int delta_z = 0, const int delta_c = 0, const unsigned int boundary_conditions = 0) const {
    if (delta_z || delta_c) throw CImgArgumentException("shift(): only x/y shifts are restated");
    if (is_empty()) return *this;
    // res(x, y) = src(x - delta_x, y - delta_y) under the boundary rule == a crop at (-delta_x, -delta_y)
    return get_crop(-delta_x, -delta_y, width() - delta_x - 1, height() - delta_y - 1, boundary_conditions);
  }

  CImg<T> get_rotate(const float angle, const unsigned int interpolation = 1, const unsigned int boundary_conditions = 0) const {
    if (is_empty()) return *this;
    CImg<T> res;
    const float nangle = cimg::mod(angle, 360.0f);
    if (boundary_conditions != 1 && cimg::mod(nangle, 90.0f) == 0) {  // orthogonal angles
      const int wm1 = width() - 1, hm1 = height() - 1;
      const int iangle = (int)nangle / 90;
      switch (iangle) {
        case 1: {
          res.assign(_height, _width, _depth, _spectrum);
          T* ptrd = res._data;
          cimg_forXYZC(res, x, y, z, c) *(ptrd++) = (*this)(y, hm1 - x, z, c);
        } break;
        case 2: {
          res.assign(_width, _height, _depth, _spectrum);
          T* ptrd = res._data;
          cimg_forXYZC(res, x, y, z, c) *(ptrd++) = (*this)(wm1 - x, hm1 - y, z, c);
        } break;
        case 3: {
          res.assign(_height, _width, _depth, _spectrum);
          T* ptrd = res._data;
          cimg_forXYZC(res, x, y, z, c) *(ptrd++) = (*this)(wm1 - y, x, z, c);
        } break;
        default:
          return *this;
      }
      return res;
    }
    if (interpolation != 1 || boundary_conditions != 3) throw CImgArgumentException("rotate(): only linear interpolation with mirror boundary is restated");
    const float rad = (float)(nangle * cimg::PI / 180.0), ca = (float)std::cos(rad), sa = (float)std::sin(rad),
                ux = cimg::abs((_width - 1) * ca), uy = cimg::abs((_width - 1) * sa), vx = cimg::abs((_height - 1) * sa),
                vy = cimg::abs((_height - 1) * ca), w2 = 0.5f * (_width - 1), h2 = 0.5f * (_height - 1);
    res.assign((int)cimg::round(1 + ux + vx), (int)cimg::round(1 + uy + vy), _depth, _spectrum);
    const float rw2 = 0.5f * (res._width - 1), rh2 = 0.5f * (res._height - 1);
    const float ww = 2.0f * width(), hh = 2.0f * height();
    cimg_forXYZC(res, x, y, z, c) {
      const float xc = x - rw2, yc = y - rh2, mx = cimg::mod(w2 + xc * ca + yc * sa, ww), my = cimg::mod(h2 - xc * sa + yc * ca, hh);
      res(x, y, z, c) = (T)_linear_atXY(mx < width() ? mx : ww - mx - 1, my < height() ? my : hh - my - 1, z, c);
    }
    return res;
  }
  CImg<T>& rotate(const float angle, const unsigned int interpolation = 1, const unsigned int boundary_conditions = 0) {
    const float nangle = cimg::mod(angle, 360.0f);
    if (nangle == 0.0f) return *this;
    return get_rotate(nangle, interpolation, boundary_conditions).move_to(*this);
  }

  // resize(sx, sy, sz, sc, interpolation, boundary): negative sizes are percentages. Interpolations restated:
  // -1/0/1 on an empty instance (allocation, WarpFields.cpp:341-342), 2 = moving average, 3 = linear.
  CImg<T> get_resize(const int size_x, const int size_y = -100, const int size_z = -100, const int size_c = -100,
                     const int interpolation_type = 1, const unsigned int boundary_conditions = 0) const {
    if (!size_x || !size_y || !size_z || !size_c) return CImg<T>();
    const unsigned int sx = (unsigned int)(size_x < 0 ? -size_x * width() / 100 : size_x), sy = (unsigned int)(size_y < 0 ? -size_y * height() / 100 : size_y),
                       sz = (unsigned int)(size_z < 0 ? -size_z * depth() / 100 : size_z), sc = (unsigned int)(size_c < 0 ? -size_c * spectrum() / 100 : size_c);
    const unsigned int nsx = sx ? sx : 1, nsy = sy ? sy : 1, nsz = sz ? sz : 1, nsc = sc ? sc : 1;
    if (nsx == _width && nsy == _height && nsz == _depth && nsc == _spectrum) return *this;
    if (is_empty()) return CImg<T>(nsx, nsy, nsz, nsc, (T)0);
    if (nsz != _depth || nsc != _spectrum) throw CImgArgumentException("resize(): only x/y resizing is restated");
    if (interpolation_type == 3) {
      CImg<T> resx = nsx == _width ? *this : (_width == 1 || _width > nsx) ? resize_axis(*this, nsx, true, _width == 1 ? 1 : 2) : resize_axis(*this, nsx, true, 3);
      return nsy == _height ? resx : (_height == 1 || _height > nsy) ? resize_axis(resx, nsy, false, _height == 1 ? 1 : 2) : resize_axis(resx, nsy, false, 3);
    }
    if (interpolation_type == 2) {  // both passes accumulate in Tfloat; stored as T once (one axis at a time is all the linear case needs)
      if (nsx != _width && nsy != _height) throw CImgArgumentException("resize(): two-axis moving average is not restated");
      return nsx != _width ? resize_axis(*this, nsx, true, 2) : resize_axis(*this, nsy, false, 2);
    }
    throw CImgArgumentException("resize(): interpolation type not restated");
  }
  CImg<T>& resize(const int size_x, const int size_y = -100, const int size_z = -100, const int size_c = -100,
                  const int interpolation_type = 1, const unsigned int boundary_conditions = 0) {
    if (!size_x || !size_y || !size_z || !size_c) return assign();
    return get_resize(size_x, size_y, size_z, size_c, interpolation_type, boundary_conditions).move_to(*this);
  }

  // permute_axes("abcd"): new axis i is old axis order[i]
  CImg<T> get_permute_axes(const char* const order) const {
    if (is_empty() || !order) return *this;
    int ax[4];
    for (int i = 0; i < 4; ++i) {
      const char ch = order[i];
      ax[i] = (ch == 'x' || ch == 'X') ? 0 : (ch == 'y' || ch == 'Y') ? 1 : (ch == 'z' || ch == 'Z') ? 2 : (ch == 'c' || ch == 'C') ? 3 : -1;
      if (ax[i] < 0) throw CImgArgumentException("permute_axes(): invalid axis order");
    }
    const unsigned int dims[4] = {_width, _height, _depth, _spectrum};
    CImg<T> res(dims[ax[0]], dims[ax[1]], dims[ax[2]], dims[ax[3]]);
    unsigned int o[4];
    cimg_forXYZC(res, x, y, z, c) {
      o[ax[0]] = x; o[ax[1]] = y; o[ax[2]] = z; o[ax[3]] = c;
      res(x, y, z, c) = (*this)(o[0], o[1], o[2], o[3]);
    }
    return res;
  }
  CImg<T>& permute_axes(const char* const order) { return get_permute_axes(order).move_to(*this); }

  // draw_image(x0, y0, sprite, mask, opacity, mask_max_value): the full-overlap case of the generic routine
  template <class ti, class tm>
  CImg<T>& draw_image(const int x0, const int y0, const CImg<ti>& sprite, const CImg<tm>& mask, const float opacity = 1, const float mask_max_value = 1) {
    if (is_empty() || !sprite || !mask) return *this;
    if (mask._width != sprite._width || mask._height != sprite._height || mask._depth != sprite._depth)
      throw CImgArgumentException("draw_image(): sprite and mask have incompatible dimensions");
    const int lX = sprite.width() - (x0 + sprite.width() > width() ? x0 + sprite.width() - width() : 0) + (x0 < 0 ? x0 : 0),
              lY = sprite.height() - (y0 + sprite.height() > height() ? y0 + sprite.height() - height() : 0) + (y0 < 0 ? y0 : 0),
              lZ = std::min(sprite.depth(), depth()), lC = std::min(sprite.spectrum(), spectrum());
    if (lX <= 0 || lY <= 0 || lZ <= 0 || lC <= 0) return *this;
    const int sx0 = x0 < 0 ? -x0 : 0, sy0 = y0 < 0 ? -y0 : 0, dx0 = x0 < 0 ? 0 : x0, dy0 = y0 < 0 ? 0 : y0;
    for (int c = 0; c < lC; ++c)
      for (int z = 0; z < lZ; ++z)
        for (int y = 0; y < lY; ++y) {
          T* ptrd = data(dx0, dy0 + y, z, c);
          const ti* ptrs = sprite.data(sx0, sy0 + y, z, c);
          const tm* ptrm = mask.data(sx0, sy0 + y, z, c % mask._spectrum);  // the mask's channels are cycled
          for (int x = 0; x < lX; ++x) {
            const float mopacity = (float)(*(ptrm++) * opacity), nopacity = cimg::abs(mopacity), copacity = mask_max_value - std::max(mopacity, 0.f);
            *ptrd = (T)((nopacity * (*(ptrs++)) + *ptrd * copacity) / mask_max_value);
            ++ptrd;
          }
        }
    return *this;
  }

  // ---- files: binary PNM only (the offline texture pools are written as P6); everything else raises like a missing codec
  CImg<T>& load(const char* const filename) {
    std::FILE* f = std::fopen(filename, "rb");
    if (!f) throw CImgIOException(std::string("load(): cannot open ") + filename);
    int magic = 0, w = 0, h = 0, maxv = 0;
    auto token = [&](int& v) -> bool {
      int ch = std::fgetc(f);
      for (;;) {
        while (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') ch = std::fgetc(f);
        if (ch == '#') { while (ch != '\n' && ch != EOF) ch = std::fgetc(f); continue; }
        break;
      }
      if (ch < '0' || ch > '9') return false;
      v = 0;
      while (ch >= '0' && ch <= '9') { v = v * 10 + (ch - '0'); ch = std::fgetc(f); }
      return true;  // the single whitespace after the token is consumed
    };
    bool ok = std::fgetc(f) == 'P';
    magic = std::fgetc(f);
    ok = ok && (magic == '5' || magic == '6') && token(w) && token(h) && token(maxv) && w > 0 && h > 0 && maxv == 255;
    if (!ok) { std::fclose(f); throw CImgIOException(std::string("load(): only 8-bit binary PNM is restated: ") + filename); }
    const int nc = magic == '6' ? 3 : 1;
    std::string raw((size_t)w * h * nc, '\0');
    const size_t got = std::fread(&raw[0], 1, raw.size(), f);
    std::fclose(f);
    if (got != raw.size()) throw CImgIOException(std::string("load(): truncated file ") + filename);
    assign(w, h, 1, nc);
    for (int c = 0; c < nc; ++c)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) (*this)(x, y, 0, c) = (T)(unsigned char)raw[((size_t)y * w + x) * nc + c];
    return *this;
  }
  const CImg<T>& save(const char* const filename) const {  // debug helper of the reference (writeMasksToFiles); P5 / P6
    std::FILE* f = std::fopen(filename, "wb");
    if (!f) throw CImgIOException(std::string("save(): cannot open ") + filename);
    const int nc = _spectrum >= 3 ? 3 : 1;
    std::fprintf(f, "P%d\n%u %u\n255\n", nc == 3 ? 6 : 5, _width, _height);
    for (int y = 0; y < height(); ++y)
      for (int x = 0; x < width(); ++x)
        for (int c = 0; c < nc; ++c) std::fputc((int)(unsigned char)(*this)(x, y, 0, c), f);
    std::fclose(f);
    return *this;
  }

 private:
  static int to_int(const float v) {
    if (!(v > -2147483648.f && v < 2147483648.f)) return (int)0x80000000;  // cvttss2si's "integer indefinite" (NaN included)
    return (int)v;
  }
  // one axis of get_resize: mode 1 = nearest (source length 1), 2 = moving average (shrinking), 3 = linear (growing)
  static CImg<T> resize_axis(const CImg<T>& s, const unsigned int n, const bool along_x, const int mode) {
    const unsigned int len = along_x ? s._width : s._height, other = along_x ? s._height : s._width;
    CImg<T> r(along_x ? n : s._width, along_x ? s._height : n, s._depth, s._spectrum);
    auto src = [&](unsigned int i, unsigned int j, unsigned int z, unsigned int c) -> const T& { return along_x ? s(i, j, z, c) : s(j, i, z, c); };
    auto dst = [&](unsigned int i, unsigned int j, unsigned int z, unsigned int c) -> T& { return along_x ? r(i, j, z, c) : r(j, i, z, c); };
    if (mode == 1) {
      cimg_forC(s, c) cimg_forZ(s, z) for (unsigned int j = 0; j < other; ++j) for (unsigned int i = 0; i < n; ++i) dst(i, j, z, c) = src((unsigned int)((double)i * len / n), j, z, c);
      return r;
    }
    if (mode == 2) {
      CImg<Tfloat> tmp(along_x ? n : s._width, along_x ? s._height : n, s._depth, s._spectrum, (Tfloat)0);
      auto acc = [&](unsigned int i, unsigned int j, unsigned int z, unsigned int c) -> Tfloat& { return along_x ? tmp(i, j, z, c) : tmp(j, i, z, c); };
      for (unsigned int a = len * n, b = len, cc = n, si = 0, t = 0; a;) {
        const unsigned int d = std::min(b, cc);
        a -= d; b -= d; cc -= d;
        cimg_forC(s, c) cimg_forZ(s, z) for (unsigned int j = 0; j < other; ++j) acc(t, j, z, c) += (Tfloat)src(si, j, z, c) * d;
        if (!b) {
          cimg_forC(s, c) cimg_forZ(s, z) for (unsigned int j = 0; j < other; ++j) acc(t, j, z, c) /= len;
          ++t;
          b = len;
        }
        if (!cc) { ++si; cc = n; }
      }
      for (size_t i = 0, m = r.size(); i < m; ++i) r._data[i] = (T)tmp._data[i];
      return r;
    }
    // linear, boundary_conditions == 0: fx = (len - 1) / (n - 1), source position accumulated in double
    const double fx = n > 1 ? (len - 1.0) / (n - 1) : 0;
    CImg<unsigned int> off(n);
    CImg<double> foff(n);
    double curr = 0, old = 0;
    for (unsigned int i = 0; i < n; ++i) {
      foff(i) = curr - (unsigned int)curr;
      old = curr;
      curr = std::min(len - 1.0, curr + fx);
      off(i) = (unsigned int)curr - (unsigned int)old;
    }
    cimg_forC(s, c) cimg_forZ(s, z) for (unsigned int j = 0; j < other; ++j) {
      unsigned int p = 0;
      for (unsigned int i = 0; i < n; ++i) {
        const double alpha = foff(i);
        const T val1 = src(p, j, z, c), val2 = p < len - 1 ? src(p + 1, j, z, c) : val1;
        dst(i, j, z, c) = (T)((1 - alpha) * val1 + alpha * val2);
        p += off(i);
      }
    }
    return r;
  }
  template <class U> friend struct CImg;
};

}  // namespace cimg_library

#endif
