// See texture_io.hpp.
#include "texture_io.hpp"

#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <zlib.h>

#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace ofdg {

namespace {

[[noreturn]] void fail(const std::string& path, const std::string& why) { throw std::runtime_error("texture " + path + ": " + why); }

std::vector<unsigned char> read_all(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("Could not open texture " + path);
  std::vector<unsigned char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  return d;
}

// interleaved RGB(A) rows -> planar B,G,R
TextureImage from_interleaved(const unsigned char* px, int w, int h, int channels, size_t row_stride, bool rows_bottom_up, bool bgr_order) {
  TextureImage t;
  t.w = w; t.h = h;
  const size_t plane = (size_t)w * h;
  t.planar_bgr.resize(3 * plane);
  for (int y = 0; y < h; ++y) {
    const unsigned char* row = px + (size_t)(rows_bottom_up ? h - 1 - y : y) * row_stride;
    unsigned char* b = t.planar_bgr.data() + (size_t)y * w;
    for (int x = 0; x < w; ++x) {
      const unsigned char* p = row + (size_t)x * channels;
      unsigned char r, g, bl;
      if (channels <= 2) { r = g = bl = p[0]; }
      else if (bgr_order) { bl = p[0]; g = p[1]; r = p[2]; }
      else { r = p[0]; g = p[1]; bl = p[2]; }
      b[x] = bl; b[plane + x] = g; b[2 * plane + x] = r;  // std::swap(c0, c2): the reference holds B,G,R planes
    }
  }
  return t;
}

TextureImage decode_ppm(const std::string& path, const std::vector<unsigned char>& d) {
  size_t i = 2;
  auto next_int = [&]() {
    for (;;) {
      if (i >= d.size()) fail(path, "truncated header");
      if (d[i] == '#') { while (i < d.size() && d[i] != '\n') ++i; }
      else if (std::isspace(d[i])) ++i;
      else break;
    }
    long v = 0;
    bool any = false;
    while (i < d.size() && std::isdigit(d[i])) { v = v * 10 + (d[i] - '0'); ++i; any = true; if (v > 1000000) break; }
    if (!any) fail(path, "bad header");
    return (int)v;
  };
  const int w = next_int(), h = next_int(), maxv = next_int();
  ++i;  // the single whitespace byte after maxval
  if (maxv != 255) fail(path, "maxval must be 255");
  if (w <= 0 || h <= 0 || w > 32768 || h > 32768) fail(path, "bad size");
  if (d.size() < i + (size_t)w * h * 3) fail(path, "truncated file");
  return from_interleaved(d.data() + i, w, h, 3, (size_t)w * 3, false, false);
}

uint32_t le32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

TextureImage decode_bmp(const std::string& path, const std::vector<unsigned char>& d) {
  if (d.size() < 54) fail(path, "truncated BMP header");
  const uint32_t data_off = le32(&d[10]), hdr = le32(&d[14]);
  const int32_t w = (int32_t)le32(&d[18]), hs = (int32_t)le32(&d[22]);
  const int bpp = d[28] | (d[29] << 8);
  const uint32_t compression = le32(&d[30]);
  if (hdr < 40) fail(path, "unsupported BMP header");
  if ((bpp != 24 && bpp != 32) || (compression != 0 && !(compression == 3 && bpp == 32))) fail(path, "only uncompressed 24/32-bit BMP is decoded");
  const int h = hs < 0 ? -hs : hs;
  if (w <= 0 || h <= 0 || w > 32768 || h > 32768) fail(path, "bad size");
  const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
  if (d.size() < data_off + stride * h) fail(path, "truncated file");
  return from_interleaved(d.data() + data_off, w, h, bpp / 8, stride, hs > 0, true);
}

TextureImage decode_png(const std::string& path, const std::vector<unsigned char>& d) {
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (d.size() < 8 || std::memcmp(d.data(), sig, 8)) fail(path, "not a PNG");
  size_t i = 8;
  int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> idat, palette;
  bool end = false;
  while (!end && i + 12 <= d.size()) {
    const uint32_t len = be32(&d[i]);
    const unsigned char* type = &d[i + 4];
    const unsigned char* body = &d[i + 8];
    if (i + 12 + (size_t)len > d.size()) fail(path, "truncated chunk");
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) fail(path, "bad IHDR");
      w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
    else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!std::memcmp(type, "IEND", 4)) end = true;
    i += 12 + (size_t)len;
  }
  if (w <= 0 || h <= 0 || w > 32768 || h > 32768) fail(path, "bad size");
  if (interlace > 1) fail(path, "bad PNG interlace method");
  const bool sub_byte = (depth == 1 || depth == 2 || depth == 4) && (ctype == 0 || ctype == 3);
  const bool wide = depth == 16 && ctype != 3;
  if (depth != 8 && !sub_byte && !wide) fail(path, "bad PNG bit depth for its colour type");
  int ch;
  switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: fail(path, "unsupported PNG colour type");
  }
  // One image (the whole picture, or one Adam7 pass of it) = ph filtered scanlines of pw pixels: undo the scanline filters
  // (on bytes; the filters' "previous pixel" is one byte back for depths below 8), then unpack sub-byte samples the way
  // libpng's expansion does (grey: scaled to 0..255; palette: the index).
  auto row_bytes = [&](int pw) { return ((size_t)pw * ch * depth + 7) / 8; };
  std::vector<unsigned char> packed;
  auto unfilter = [&](const unsigned char* raw, int pw, int ph, unsigned char* dst) {
    const size_t stride = row_bytes(pw);
    const size_t fb = (size_t)std::max(1, ch * depth / 8);
    unsigned char* base = dst;
    if (sub_byte || wide) { packed.assign(stride * ph, 0); base = packed.data(); }
    for (int y = 0; y < ph; ++y) {
      const unsigned char* in = raw + (size_t)y * (stride + 1);
      unsigned char* cur = base + (size_t)y * stride;
      const unsigned char* up = y ? cur - stride : nullptr;
      const int ft = in[0];
      for (size_t x = 0; x < stride; ++x) {
        const int a = x >= fb ? cur[x - fb] : 0, b = up ? up[x] : 0, c = (up && x >= fb) ? up[x - fb] : 0;
        int pred;
        switch (ft) {
          case 0: pred = 0; break;
          case 1: pred = a; break;
          case 2: pred = b; break;
          case 3: pred = (a + b) >> 1; break;
          case 4: {
            const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
            pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            break;
          }
          default: fail(path, "bad PNG filter");
        }
        cur[x] = (unsigned char)(in[1 + x] + pred);
      }
    }
    if (wide) {
      // 16-bit samples: the reference reads them into a CImg<unsigned char>, i.e. CImg's loader casts each 16-bit value to
      // unsigned char -- its LOW byte (big-endian in the file: the second one). Reproduced as it is.
      for (int y = 0; y < ph; ++y)
        for (size_t x = 0; x < (size_t)pw * ch; ++x) dst[(size_t)y * pw * ch + x] = packed[(size_t)y * stride + 2 * x + 1];
    }
    if (sub_byte) {
      const int scale = ctype == 0 ? 255 / ((1 << depth) - 1) : 1;
      for (int y = 0; y < ph; ++y)
        for (int x = 0; x < pw; ++x) {
          const unsigned char byte = packed[(size_t)y * stride + ((size_t)x * depth) / 8];
          const int shift = 8 - depth - (int)(((size_t)x * depth) % 8);
          dst[(size_t)y * pw + x] = (unsigned char)(((byte >> shift) & ((1 << depth) - 1)) * scale);
        }
    }
  };
  const size_t stride = (size_t)w * ch;
  std::vector<unsigned char> px(stride * h);
  if (!interlace) {
    std::vector<unsigned char> raw((row_bytes(w) + 1) * h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) fail(path, "inflate failed");
    unfilter(raw.data(), w, h, px.data());
  } else {  // Adam7: seven reduced images one after the other in the stream, each filtered on its own
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    int pw[7], ph[7];
    size_t total = 0;
    for (int k = 0; k < 7; ++k) {
      pw[k] = (w - xs[k] + dx[k] - 1) / dx[k];
      ph[k] = (h - ys[k] + dy[k] - 1) / dy[k];
      if (pw[k] > 0 && ph[k] > 0) total += (row_bytes(pw[k]) + 1) * ph[k];
    }
    std::vector<unsigned char> raw(total);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) fail(path, "inflate failed");
    std::vector<unsigned char> pass;
    size_t off = 0;
    for (int k = 0; k < 7; ++k) {
      if (pw[k] <= 0 || ph[k] <= 0) continue;
      pass.assign((size_t)pw[k] * ch * ph[k], 0);
      unfilter(raw.data() + off, pw[k], ph[k], pass.data());
      off += (row_bytes(pw[k]) + 1) * ph[k];
      for (int j = 0; j < ph[k]; ++j)
        for (int i2 = 0; i2 < pw[k]; ++i2)
          std::memcpy(&px[((size_t)(ys[k] + j * dy[k]) * w + xs[k] + (size_t)i2 * dx[k]) * ch], &pass[((size_t)j * pw[k] + i2) * ch], (size_t)ch);
    }
  }
  if (ctype == 3) {  // palette -> RGB
    if (palette.size() < 3) fail(path, "PNG palette missing");
    std::vector<unsigned char> rgb((size_t)w * h * 3);
    for (size_t k = 0; k < (size_t)w * h; ++k) {
      const size_t e = (size_t)px[k] * 3;
      if (e + 2 >= palette.size()) fail(path, "PNG palette index out of range");
      rgb[3 * k] = palette[e]; rgb[3 * k + 1] = palette[e + 1]; rgb[3 * k + 2] = palette[e + 2];
    }
    return from_interleaved(rgb.data(), w, h, 3, (size_t)w * 3, false, false);
  }
  return from_interleaved(px.data(), w, h, ch, stride, false, false);
}

// JPEG (baseline and progressive) through nvJPEG, the decoder that ships with the CUDA toolkit: straight to planar B,G,R on
// the current device, then to the host. The reference decodes JPEG through CImg -> libjpeg; decoders agree on the bitstream
// but may round the inverse DCT / chroma upsampling differently by an LSB -- a property of the texture database, not of the
// render path (whose parity is defined on the pixels the pool holds).
TextureImage decode_jpeg(const std::string& path, const std::vector<unsigned char>& d) {
  struct Nv {
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t st = nullptr;
    ~Nv() {
      if (st) nvjpegJpegStateDestroy(st);
      if (h) nvjpegDestroy(h);
    }
  } nv;
  if (nvjpegCreateSimple(&nv.h) != NVJPEG_STATUS_SUCCESS) fail(path, "nvjpegCreateSimple failed (JPEG textures need a CUDA device)");
  if (nvjpegJpegStateCreate(nv.h, &nv.st) != NVJPEG_STATUS_SUCCESS) fail(path, "nvjpegJpegStateCreate failed");
  int comps = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
  nvjpegChromaSubsampling_t sub;
  if (nvjpegGetImageInfo(nv.h, d.data(), d.size(), &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) fail(path, "not a JPEG image nvJPEG can parse");
  const int w = ws[0], h = hs[0];
  if (w <= 0 || h <= 0) fail(path, "bad JPEG size");
  const size_t plane = (size_t)w * h;
  unsigned char* dev = nullptr;
  if (cudaMalloc(&dev, 3 * plane) != cudaSuccess) fail(path, "cudaMalloc failed while decoding a JPEG");
  nvjpegImage_t out{};
  for (int c = 0; c < 3; ++c) { out.channel[c] = dev + (size_t)c * plane; out.pitch[c] = (size_t)w; }
  const nvjpegStatus_t rc = nvjpegDecode(nv.h, nv.st, d.data(), d.size(), NVJPEG_OUTPUT_BGR, &out, nullptr);  // planar B, G, R
  TextureImage t;
  t.w = w; t.h = h;
  t.planar_bgr.resize(3 * plane);
  const cudaError_t ce = rc == NVJPEG_STATUS_SUCCESS ? cudaMemcpy(t.planar_bgr.data(), dev, 3 * plane, cudaMemcpyDeviceToHost) : cudaSuccess;
  cudaFree(dev);
  if (rc != NVJPEG_STATUS_SUCCESS) fail(path, "nvjpegDecode failed (status " + std::to_string((int)rc) + ")");
  if (ce != cudaSuccess) fail(path, "copying the decoded JPEG to the host failed");
  return t;
}

}  // namespace

TextureImage load_texture_file(const std::string& path) {
  const std::vector<unsigned char> d = read_all(path);
  if (d.size() >= 2 && d[0] == 'P' && d[1] == '6') return decode_ppm(path, d);
  if (d.size() >= 2 && d[0] == 'B' && d[1] == 'M') return decode_bmp(path, d);
  if (d.size() >= 4 && d[0] == 0x89 && d[1] == 'P' && d[2] == 'N' && d[3] == 'G') return decode_png(path, d);
  if (d.size() >= 3 && d[0] == 0xFF && d[1] == 0xD8 && d[2] == 0xFF) return decode_jpeg(path, d);
  fail(path, "unsupported image format (decoders: binary PPM, uncompressed BMP, PNG up to 8 bits per sample incl. Adam7, JPEG)");
}

std::vector<std::string> read_texture_list(const std::string& listfile) {
  std::ifstream infile(listfile);
  if (infile.bad() || !infile.is_open()) throw std::runtime_error("Could not open texture collection");  // DataGenerator.cpp:121
  // The reference's loop, DataGenerator.cpp:123-126:  while (!eof) { getline(path); if (eof) break; load(path); }
  // A last line that is not terminated by a newline sets eof inside getline and is therefore NOT loaded; pool slot k is
  // tex_id % pool size, so the pool must have exactly the reference's size for the k-th task to pick the same texture.
  // An empty line would make CImg::load("") throw in the reference; here it is an error as well.
  std::vector<std::string> paths;
  std::string path;
  while (!infile.eof()) {
    std::getline(infile, path);
    if (infile.eof()) break;
    while (!path.empty() && path.back() == '\r') path.pop_back();  // (lists written on Windows)
    if (path.empty()) throw std::runtime_error("texture collection " + listfile + ": empty line (the reference's CImg::load(\"\") fails there too)");
    paths.push_back(path);
  }
  if (paths.empty()) throw std::runtime_error("texture collection is empty (note: a last line without a trailing newline is not read, like the reference)");
  return paths;
}

}  // namespace ofdg
