// GPU warp-field producer (mode 9); see warpfields.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ofdg {

struct WfDisplacer {  // one Displacer + its Gaussian2D support (WarpFields.cpp:88-112, 191-260)
  int kind;           // 0 translation, 1 rotation, 2 zoom
  float cx, cy, dx, dy, sin_o, cos_o, sin_no, cos_no, factor, ifactor;
  float scx, scy, a, b, c, d, ratio_x_y, sigma_sq, gauss_prefactor, normalizer;
};

// Work buffers of wf_generate for one canvas size; kept by callers that produce fields again and again
// (cudaMalloc / cudaFree per call would synchronise the device each time).
struct WfScratch {
  float *flow = nullptr, *iflow = nullptr, *tmp = nullptr;
  unsigned char* flagged = nullptr;
  WfDisplacer* d_ds = nullptr;
  int S = 0;
  void reserve(int canvas);
  void release();
};

// Writes n_fields crops [n][flow|iflow][channel][H+1][W+1] to d_out (device). Returns the number of launches.
// scratch == nullptr: temporary work buffers (allocated and freed inside the call).
int wf_generate(int W, int H, uint32_t seed, int n_fields, float* d_out, cudaStream_t s, WfScratch* scratch = nullptr);
// reach[i] = ceil(max |iflow| over the finite values of crop i), clamped to 1e6: how far the inverse field can move a mask.
int wf_reach(int W, int H, const float* d_fields, int n_fields, int* d_reach, cudaStream_t s);

}  // namespace ofdg
