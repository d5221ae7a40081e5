// GPU producer of the mode-9 non-rigid warp fields: replaces the reference's
// WarpFields::CropGenerator (10 CPU threads, /root/reference/src/caffe/WarpFields.cpp:469-641).
//   host   draws the 9x7 hex grid of random displacers + rotated-Gaussian supports   (WF.cpp:570-610)
//   device samples the elementary forward / inverse fields on the 3*max(W,H) canvas    (WF.cpp:347-354)
//          17 ping-pong self-compositions each, out-of-bounds flagging -> NaN          (WF.cpp:366-434)
//          clamp_near_zeros(1e-3) and the 8x5 crops of (W+1)x(H+1)                     (WF.cpp:444-455, 619-634)
// The reference seeds this from std::random_device; here the seed is explicit. Values agree with the
// CPU restatement (oracle/warpfields.cpp) up to the device's expf.
#include "warpfields.cuh"

#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

namespace ofdg {

namespace {

__device__ __forceinline__ float support_at(const WfDisplacer& d, float x, float y) {  // Gaussian2D::at, WF.cpp:101-112
  const float rx = d.a * (x - d.scx) + d.b * (y - d.scy);
  const float ry = (d.c * (x - d.scx) + d.d * (y - d.scy)) * d.ratio_x_y;
  const float dist_sq = rx * rx + ry * ry;
  return d.normalizer * (d.gauss_prefactor * expf(-dist_sq / (2 * d.sigma_sq)));
}

__global__ void wf_elementary_kernel(const WfDisplacer* ds, int n, int S, float* flow, float* iflow) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, P = (size_t)S * S;
  if (i >= P) return;
  const float x = (float)(i % S), y = (float)(i / S);
  float fx = 0, fy = 0, ix = 0, iy = 0;
  for (int k = 0; k < n; ++k) {
    const WfDisplacer d = ds[k];
    const float w = support_at(d, x, y);
    float ax, ay, bx, by;
    if (d.kind == 0) { ax = d.dx; ay = d.dy; bx = -d.dx; by = -d.dy; }
    else {
      const float ddx = x - d.cx, ddy = y - d.cy;
      if (d.kind == 1) {
        ax = (d.cos_no * ddx - d.sin_no * ddy) - ddx; ay = (d.sin_no * ddx + d.cos_no * ddy) - ddy;
        bx = (d.cos_o * ddx - d.sin_o * ddy) - ddx; by = (d.sin_o * ddx + d.cos_o * ddy) - ddy;
      } else {
        ax = d.factor * ddx - ddx; ay = d.factor * ddy - ddy;
        bx = d.ifactor * ddx - ddx; by = d.ifactor * ddy - ddy;
      }
    }
    fx += ax * w; fy += ay * w; ix += bx * w; iy += by * w;
  }
  flow[i] = fx; flow[P + i] = fy; iflow[i] = ix; iflow[P + i] = iy;
}

__device__ __forceinline__ float neumann(const float* f, int S, float fx, float fy) {  // CImg _linear_atXY
  const float nfx = fx <= 0 ? 0 : (fx >= S - 1 ? (float)(S - 1) : fx), nfy = fy <= 0 ? 0 : (fy >= S - 1 ? (float)(S - 1) : fy);
  const unsigned x = (unsigned)nfx, y = (unsigned)nfy;
  const float dx = nfx - x, dy = nfy - y;
  const unsigned nx = dx > 0 ? x + 1 : x, ny = dy > 0 ? y + 1 : y;
  const float Icc = f[(size_t)y * S + x], Inc = f[(size_t)y * S + nx], Icn = f[(size_t)ny * S + x], Inn = f[(size_t)ny * S + nx];
  return Icc + dx * (Inc - Icc + dy * (Icc + Inn - Icn - Inc)) + dy * (Icn - Icc);
}

__global__ void wf_compose_kernel(const float* from, float* to, unsigned char* flagged, int S) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, P = (size_t)S * S;
  if (i >= P) return;
  const int x = (int)(i % S), y = (int)(i / S);
  const float fx = from[i], fy = from[P + i];
  if (x + fx < 0 || x + fx >= S || y + fy < 0 || y + fy >= S) {
    flagged[i] = 255;
    to[i] = fx; to[P + i] = fy;
    return;
  }
  to[i] = fx + neumann(from, S, x + fx, y + fy);
  to[P + i] = fy + neumann(from + P, S, x + fx, y + fy);
}

__global__ void wf_finish_kernel(float* field, const unsigned char* flagged, int S) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, P = (size_t)S * S;
  if (i >= P) return;
  const int x = (int)(i % S), y = (int)(i / S);
  float fx = field[i], fy = field[P + i];
  const bool out = (x + fx < 0 || x + fx >= S || y + fy < 0 || y + fy >= S) || flagged[i];
  if (out) { fx = nanf(""); fy = fx; }
  if (fabsf(fx) < 1e-3f) fx = 0.f;  // clamp_near_zeros (NaN compares false and stays)
  if (fabsf(fy) < 1e-3f) fy = 0.f;
  field[i] = fx; field[P + i] = fy;
}

__global__ void wf_crop_kernel(const float* field, int S, int x0, int y0, int W1, int H1, float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)W1 * H1;
  if (i >= 2 * n) return;
  const int c = (int)(i / n), r = (int)(i % n), xx = r % W1, yy = r / W1;
  out[i] = field[(size_t)c * S * S + (size_t)(y0 + yy) * S + x0 + xx];
}

}  // namespace

// CropGenerator::worker_thread_loop's random scene (WF.cpp:570-610), std::mt19937(seed) instead of random_device
void wf_draw_displacers(std::mt19937& mersenne, int big_size, std::vector<WfDisplacer>& ds) {
  std::uniform_int_distribution<> displacer_type(0, 2);
  std::uniform_real_distribution<> generic_param(-1, 1);
  ds.clear();
  const int spacing{200};
  const int isosceles_spacing{(int)(spacing / 2. * std::sqrt(3.))};
  const int rows{(big_size + isosceles_spacing - 1) / isosceles_spacing};
  const int cols{big_size / spacing};
  for (int yidx = 0; yidx < rows; ++yidx)
    for (int xidx = 0; xidx < cols; ++xidx) {
      const int x = xidx * spacing + (yidx % 2 == 1 ? spacing / 2 : 0) + spacing / 2;
      const int y = yidx * isosceles_spacing + spacing / 2;
      WfDisplacer d{};
      d.kind = displacer_type(mersenne);
      auto g = [&]() { return generic_param(mersenne); };
      d.factor = 1; d.ifactor = 1;
      if (d.kind == 0) { d.dx = (float)(g() * 3e-4); d.dy = (float)(g() * 3e-4); }
      else if (d.kind == 1) {
        d.cx = (float)(x + g() * 10); d.cy = (float)(y + g() * 10);
        const float omega = (float)(g() * M_PI * 2e-6);
        d.sin_o = std::sin(omega); d.cos_o = std::cos(omega); d.sin_no = std::sin(-omega); d.cos_no = std::cos(-omega);
      } else {
        d.cx = (float)(x + g() * 10); d.cy = (float)(y + g() * 10);
        d.factor = (float)(1 + g() * 2e-6); d.ifactor = (float)(1. / d.factor);
      }
      // Supports::Gaussian2D (WF.cpp:88-99)
      d.scx = (float)(x + g() * 10); d.scy = (float)(y + g() * 10);
      const float sigma_x = (float)(50 + g() * 20), sigma_y = (float)(50 + g() * 20), angle = (float)(g() * M_PI);
      d.a = std::cos(angle); d.b = -std::sin(angle); d.c = std::sin(angle); d.d = std::cos(angle);
      d.ratio_x_y = sigma_x / sigma_y;
      d.sigma_sq = sigma_x * sigma_x;
      d.gauss_prefactor = (float)(1 / std::sqrt(2 * M_PI * d.sigma_sq));
      d.normalizer = 1.f / (d.gauss_prefactor * std::exp(-0.f / (2 * d.sigma_sq)));  // 1 / raw_at(cx, cy)
      ds.push_back(d);
    }
}

void WfScratch::reserve(int canvas) {
  if (canvas <= S) return;
  release();
  const size_t P = (size_t)canvas * canvas;
  cudaMalloc(&flow, 2 * P * sizeof(float)); cudaMalloc(&iflow, 2 * P * sizeof(float)); cudaMalloc(&tmp, 2 * P * sizeof(float));
  cudaMalloc(&flagged, P); cudaMalloc(&d_ds, 256 * sizeof(WfDisplacer));
  S = canvas;
}
void WfScratch::release() {
  cudaFree(flow); cudaFree(iflow); cudaFree(tmp); cudaFree(flagged); cudaFree(d_ds);
  flow = iflow = tmp = nullptr; flagged = nullptr; d_ds = nullptr; S = 0;
}

__global__ void wf_reach_kernel(const float* fields, int n2, size_t per_field, int* reach) {
  // one block per crop: max |v| over the finite values of its inverse field (the second half of the crop's record)
  __shared__ float s_max[256];
  const float* ifl = fields + (size_t)blockIdx.x * per_field + n2;
  float mx = 0.f;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const float v = fabsf(ifl[i]);
    if (v <= 3.0e38f) mx = fmaxf(mx, v);  // (NaN and inf fail the comparison)
  }
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + d]);
    __syncthreads();
  }
  if (threadIdx.x == 0) reach[blockIdx.x] = (int)ceilf(fminf(s_max[0], 1.0e6f));
}
int wf_reach(int W, int H, const float* d_fields, int n_fields, int* d_reach, cudaStream_t s) {
  const int n2 = 2 * (W + 1) * (H + 1);
  wf_reach_kernel<<<n_fields, 256, 0, s>>>(d_fields, n2, (size_t)2 * n2, d_reach);
  return 1;
}

int wf_generate(int W, int H, uint32_t seed, int n_fields, float* d_out, cudaStream_t s, WfScratch* scratch) {
  const int S = std::max(W, H) * 3, W1 = W + 1, H1 = H + 1;
  const size_t P = (size_t)S * S, per_field = (size_t)2 * 2 * W1 * H1;
  WfScratch own;
  WfScratch& sc = scratch ? *scratch : own;
  sc.reserve(S);
  float *flow = sc.flow, *iflow = sc.iflow, *tmp = sc.tmp;
  unsigned char* flagged = sc.flagged;
  WfDisplacer* d_ds = sc.d_ds;
  int launches = 0;
  std::mt19937 mersenne(seed);
  std::vector<WfDisplacer> ds;
  const int tb = 256, gb = (int)((P + tb - 1) / tb);
  int produced = 0;
  while (produced < n_fields) {
    wf_draw_displacers(mersenne, S, ds);
    cudaMemcpyAsync(d_ds, ds.data(), ds.size() * sizeof(WfDisplacer), cudaMemcpyHostToDevice, s);
    wf_elementary_kernel<<<gb, tb, 0, s>>>(d_ds, (int)ds.size(), S, flow, iflow);
    ++launches;
    float* fields[2] = {flow, iflow};
    for (int k = 0; k < 2; ++k) {
      cudaMemcpyAsync(tmp, fields[k], 2 * P * sizeof(float), cudaMemcpyDeviceToDevice, s);
      cudaMemsetAsync(flagged, 0, P, s);
      for (int iter = 17; iter > 0; --iter) {
        const float* from = (iter % 2 == 1) ? tmp : fields[k];
        float* to = (iter % 2 == 1) ? fields[k] : tmp;
        wf_compose_kernel<<<gb, tb, 0, s>>>(from, to, flagged, S);
        ++launches;
      }
      wf_finish_kernel<<<gb, tb, 0, s>>>(fields[k], flagged, S);
      ++launches;
    }
    for (int y = H / 4; y < S - 5 * H / 4 && produced < n_fields; y += H / 3)
      for (int x = W / 4; x < S - 5 * W / 4 && produced < n_fields; x += W / 3) {
        float* dst = d_out + (size_t)produced * per_field;
        const int n2 = 2 * W1 * H1;
        wf_crop_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(flow, S, x, y, W1, H1, dst);
        wf_crop_kernel<<<(n2 + 255) / 256, 256, 0, s>>>(iflow, S, x, y, W1, H1, dst + n2);
        launches += 2;
        ++produced;
      }
    cudaStreamSynchronize(s);  // the host redraws `ds` for the next canvas
  }
  if (!scratch) own.release();
  return launches;
}

}  // namespace ofdg
