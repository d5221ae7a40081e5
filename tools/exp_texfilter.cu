// Experiment: how close is the texture unit's bilinear filter (uchar4, normalised-float read, 8-bit weights) to
// agg::span_image_filter_rgb_bilinear ((32768 + sum w_k S_k) >> 16 with w from the 8-bit fractions fx, fy)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/exp_texfilter tools/exp_texfilter.cu && /tmp/exp_texfilter
// Prints, over every (fx, fy) in [0,255]^2 at a few thousand texel positions of a random image (plus a high-contrast one):
// the number of channel values that differ from AGG's and the largest difference, for two float -> byte conversions.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void probe(cudaTextureObject_t tex, const uchar4* img, int w, int h, int n_pos, unsigned long long* stats, float* raw_out) {
  // block = one position, threads over (fx, fy)
  const int pos = blockIdx.x;
  const int xl = 1 + (pos * 37) % (w - 3), yl = 1 + (pos * 101) % (h - 3);
  unsigned long long diff_a = 0, diff_b = 0, max_a = 0, max_b = 0, diff_raw = 0;
  for (int t = threadIdx.x; t < 65536; t += blockDim.x) {
    const int fx = t & 255, fy = t >> 8;
    const int x_hr = (xl << 8) | fx, y_hr = (yl << 8) | fy;
    const float u = (float)x_hr * (1.f / 256.f) + 0.5f, v = (float)y_hr * (1.f / 256.f) + 0.5f;
    const float4 r = tex2D<float4>(tex, u, v);
    const uchar4 p00 = img[yl * w + xl], p10 = img[yl * w + xl + 1], p01 = img[(yl + 1) * w + xl], p11 = img[(yl + 1) * w + xl + 1];
    const unsigned w00 = (256 - fx) * (256 - fy), w10 = fx * (256 - fy), w01 = (256 - fx) * fy, w11 = fx * fy;
    const unsigned s[3] = {w00 * p00.x + w10 * p10.x + w01 * p01.x + w11 * p11.x, w00 * p00.y + w10 * p10.y + w01 * p01.y + w11 * p11.y,
                           w00 * p00.z + w10 * p10.z + w01 * p01.z + w11 * p11.z};
    const float rr[3] = {r.x, r.y, r.z};
    for (int c = 0; c < 3; ++c) {
      const int agg = (int)((32768u + s[c]) >> 16);
      const int a = (int)(rr[c] * 255.f + 0.5f);                 // round to nearest
      const int b = (int)__fmaf_rn(rr[c], 255.f, 0.5f);          // same with one rounding
      const unsigned long long da = (unsigned long long)abs(a - agg), db = (unsigned long long)abs(b - agg);
      diff_a += da != 0; diff_b += db != 0;
      if (da > max_a) max_a = da;
      if (db > max_b) max_b = db;
      // is the raw filter output the exact sum / (65536 * 255)?
      const float exact = (float)((double)s[c] / (65536.0 * 255.0));
      diff_raw += exact != rr[c];
      if (pos == 0 && c == 0 && raw_out) raw_out[t] = rr[c] * 255.f - (float)((double)s[c] / 65536.0);
    }
  }
  atomicAdd(&stats[0], diff_a); atomicAdd(&stats[1], diff_b); atomicMax(&stats[2], max_a); atomicMax(&stats[3], max_b); atomicAdd(&stats[4], diff_raw);
}

int main() {
  const int w = 256, h = 64, n_pos = 2048;
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<uchar4> img(w * h);
    srand(7 + variant);
    for (auto& p : img) {
      if (variant == 0) p = make_uchar4(rand() & 255, rand() & 255, rand() & 255, 0);
      else p = make_uchar4((rand() & 1) * 255, (rand() & 1) ? 254 : 1, rand() & 255, 0);  // extreme contrast
    }
    uchar4* d_img;
    size_t pitch;
    CK(cudaMallocPitch(&d_img, &pitch, w * sizeof(uchar4), h));
    CK(cudaMemcpy2D(d_img, pitch, img.data(), w * sizeof(uchar4), w * sizeof(uchar4), h, cudaMemcpyHostToDevice));
    uchar4* d_lin;
    CK(cudaMalloc(&d_lin, img.size() * sizeof(uchar4)));
    CK(cudaMemcpy(d_lin, img.data(), img.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d_img;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    unsigned long long* d_stats;
    CK(cudaMalloc(&d_stats, 5 * sizeof(unsigned long long)));
    CK(cudaMemset(d_stats, 0, 5 * sizeof(unsigned long long)));
    float* d_raw;
    CK(cudaMalloc(&d_raw, 65536 * sizeof(float)));
    probe<<<n_pos, 256>>>(tex, d_lin, w, h, n_pos, d_stats, d_raw);
    CK(cudaDeviceSynchronize());
    unsigned long long st[5];
    CK(cudaMemcpy(st, d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    std::vector<float> raw(65536);
    CK(cudaMemcpy(raw.data(), d_raw, 65536 * sizeof(float), cudaMemcpyDeviceToHost));
    float worst = 0;
    for (float v : raw) worst = fabsf(v) > worst ? fabsf(v) : worst;
    const double total = 3.0 * 65536.0 * n_pos;
    printf("variant %d (pitch %zu): values %.0f | round-nearest: %llu differ (%.4f %%), max |diff| %llu | fma form: %llu differ, max %llu | raw != exact: %llu (%.2f %%), worst |255 r - exact| at position 0: %g\n",
           variant, pitch, total, st[0], 100.0 * st[0] / total, st[2], st[1], st[3], st[4], 100.0 * st[4] / total, worst);
    cudaDestroyTextureObject(tex);
    cudaFree(d_img); cudaFree(d_lin); cudaFree(d_stats); cudaFree(d_raw);
  }
  return 0;
}
