// Kernel-side parameter block and launch entry points of the sm_100a render path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "flat_scene.h"

namespace ofdg {

// Everything the kernels of one render call need (passed by value).
struct RenderArgs {
  const FlatSample* samples;
  const FlatObject* objects;
  const FlatShape* shapes;
  const FlatVertex* verts;
  int batch;
  int W, H;          // output size
  int use_aa;
  // texture pool: RGBX8 interleaved, textures of any size back to back, described by tex_info[n]
  const uchar4* pool;
  const TexInfo* tex_info;
  // per-sample prepared background texture, RGBX8 [batch][2H][2W]
  uchar4* bg;
  // CImg linear-resize tables for every source length: row `len` of pos_x/alpha_x ([2W][2W]) describes len -> 2W
  const int* pos_x;
  const double* alpha_x;
  const int* pos_y;        // [2H][2H]
  const double* alpha_y;
  // per-sample, per-tile lists of the objects whose boxes touch the tile (bin_kernel -> render_kernel)
  uint8_t* tile_hits;         // [batch][tiles][TILE_HIT_STRIDE]: byte 0 = count (255: too many, rescan), then object indices in z-order
  // split render path: (object, tile) pairs (bin_pairs_kernel -> raster_pairs_kernel -> shade_kernel)
  int2* tile_range;           // [batch][tiles] {first pair, pair count}, pairs of a tile in z-order
  int4* pair_list;            // [pair_cap] {sample * 256 + object, tile, first outline (absolute), outline count | composite << 16}
  uint32_t float_bias;        // 0x4B000000 (2^23 as float bits), handed in as a parameter: the byte -> float permutes that insert it keep immediate selectors
  int prep_w, prep_h;         // extent of the largest needed part of a prepared background (0: the whole canvas): bg_prep_kernel's grid
  uint32_t* pair_masks;       // [pair_cap][AA 0 | AA 1 | non-AA 0 | non-AA 1][TH][32] four pixels per word
  // span-interpolator rows (agg::span_interpolator_linear::begin, DataGenerator.cpp:203-221) hoisted out of the shade kernel:
  // two int4 per row = {x1, lft_x, rem_x, y1} {lft_y, rem_y, -, -} of dda2_line_interpolator over the row's span
  int4* bg_rows;              // [batch][H][2] background rows (written by bin_pairs_kernel), span length 2W
  int4* pair_rows;            // [pair_cap][TH][2] the pair's object over the tile's rows (written by raster_pairs_kernel), span length W
  int* pair_ctl;              // [0] pairs claimed (atomic), [1] set when they exceed pair_cap (cannot happen: pair_cap is an upper bound; the kernels then skip the batch), [2] the raster kernel's work queue
  int pair_cap;
  int* pair_overflow;         // mapped host int, raised together with pair_ctl[1] (the host reports it after its next synchronisation)
  const uint8_t* comp_lut;    // [2][256][256] MovingObjectComposite::renderMasks' rules tabulated (composite_lut_kernel): additive, then subtractive; index u * 256 + v
  // mode 9
  const float* fields;        // [n][flow|iflow][channel][H+1][W+1]
  int n_fields;
  int n_deform;               // outlines whose frame-1 masks are warped (scratch slots)
  const int* deform_shape;    // [n_deform] index into shapes
  const int* deform_field;    // [n_deform] field id
  uint8_t* mask_raw;          // [n_deform][AA|noAA][H][W] frame-1 masks before the warp
  uint8_t* mask_warp;         // [n_deform][AA|noAA][H][W] after the warp
  const int* fpos_x;          // CImg linear-resize tables (W+1 -> 2W, H+1 -> 2H) for the background's fields
  const double* falpha_x;
  const int* fpos_y;
  const double* falpha_y;
  // outputs (device): NCHW float blobs
  float* img0;
  float* img1;
  float* flow;
  // extra tops (device, each may be null): what the reference's RenderCore holds besides the three blobs
  float* flow_bw;       // [batch][2][H][W] computeFlowImage(inverse = true): flow of frame 1's pixels back to frame 0
  float* top_id0;       // [batch][1][H][W] index_image0 as float (1 = background, 10 + k = k-th foreground object)
  float* top_id1;
  float* occlusion;     // [batch][1][H][W] 1 where frame 0's pixel is not visible in frame 1 (include/ofdg/ofdg.h)
  uint8_t* ids8;        // [batch][2][H][W] scratch: per-sample object ranks (0 = background, k + 1), feeds the occlusion pass
  // parity instrumentation (device, may be null)
  uint8_t* dbg_masks;   // [batch][max_objs][4][H][W]
  int dbg_max_objs;
  uint32_t* dbg_id0;    // [batch][H][W]
  uint32_t* dbg_id1;
  uint8_t* frames8;     // [batch][2][3][H][W] the frames as bytes: uint8 transport of the host-blob path (img0/img1 may then be null) and parity checks
};

constexpr int TILE_HIT_STRIDE = 32;  // 1 count byte + up to 31 object indices per tile

// Launchers; each returns the number of kernels it launched.
int launch_bin(const RenderArgs& a, cudaStream_t s);
size_t tile_hits_bytes(int batch, int W, int H);
int launch_background_prep(const RenderArgs& a, cudaStream_t s);
int launch_render(const RenderArgs& a, cudaStream_t s);  // + the occlusion pass when a.occlusion is set
// Split path: bin + raster + shade (+ the fused kernel, which only does work when the pair buffer overflowed).
// Replaces launch_bin + launch_render; a.pair_* must be set.
// before_shade, if given, is recorded just before the shade kernel. done: how much of the step the caller has already queued
// for this batch on a forked stream and made s wait for -- 1: launch_bin_pairs (one block per sample, a 12 us latency-bound
// launch, beside the background preparation), 2: launch_raster_pairs as well (the masks need the scene and the pair list only).
int launch_bin_pairs(const RenderArgs& a, cudaStream_t s);
int launch_raster_pairs(const RenderArgs& a, cudaStream_t s);
int launch_render_split(const RenderArgs& a, cudaStream_t s, cudaEvent_t before_shade = nullptr, int done = 0);
size_t pair_mask_bytes_per_pair();
size_t pair_row_bytes_per_pair();
int launch_deform_prepass(const RenderArgs& a, cudaStream_t s);  // mode 9 only; no-op when n_deform == 0

// Foreground view (W x H) of a texture smaller than W x H, CImg linear resize; tmp holds W x h pixels, pos / alpha max(W, H) entries.
int launch_fg_resize(const uchar4* tex, int w, int h, int W, int H, uchar4* out, uint32_t* tmp, int* pos, double* alpha, cudaStream_t s);
void launch_resize_tables(int* pos, double* alpha, int n, cudaStream_t s);  // one-time, all lengths 1..n-1 -> n
// Scene upload without the copy engines: one kernel pulls the flattened arrays out of mapped pinned host memory.
// (In the host-blob pipeline the copy engine is busy with large device-to-host copies; a small host-to-device
// memcpy queued behind one of them stalls the compute stream for the length of that copy.)
struct UploadSegments {
  static constexpr int kMax = 6;
  const void* src[kMax];  // device-visible address of the pinned source, 16-byte aligned
  void* dst[kMax];        // 16-byte aligned, capacity rounded up to 16 bytes
  unsigned n16[kMax];     // length in 16-byte units
  int n;
};
int launch_scene_upload(const UploadSegments& u, cudaStream_t s);
void launch_planar_to_rgbx(const uint8_t* planar, uchar4* out, int n, int w, int h, cudaStream_t s);
void launch_rgbx_to_planar(const uchar4* in, uint8_t* planar, int w, int h, cudaStream_t s);
void launch_synth_textures(uchar4* out, int n, int w, int h, uint64_t seed, int first_index, cudaStream_t s);
void launch_composite_luts(uint8_t* add_lut, uint8_t* sub_lut, cudaStream_t s);
void launch_bg_to_planar(const uchar4* bg, uint8_t* planar, int batch, int w2, int h2, cudaStream_t s);

}  // namespace ofdg
