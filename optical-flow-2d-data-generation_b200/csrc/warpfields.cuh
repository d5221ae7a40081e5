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

// Writes n_fields crops [n][flow|iflow][channel][H+1][W+1] to d_out (device). Returns the number of launches.
int wf_generate(int W, int H, uint32_t seed, int n_fields, float* d_out, cudaStream_t s);

}  // namespace ofdg
