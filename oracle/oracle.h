/*
 * oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This is a CPU restatement of the reference generator's
 * render path (see oracle.cpp for the file:line map). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (optical-flow-2d-data-generation_b200/) never does.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors, and its two
 * arithmetic dependencies (AGG 2.4, CImg) plus Caffe are not vendored, so the reference
 * cannot be built in this environment. The oracle is pinned only by known-answer tests
 * derived by hand from the published AGG algorithm (tests/test_oracle_kat.py).
 */
#ifndef OFDG_ORACLE_H_
#define OFDG_ORACLE_H_

#include <stdint.h>

#include "ofdg/scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_config {
  int32_t W, H;                 /* output size (DGEN_WIDTH x DGEN_HEIGHT) */
  int32_t mode;                 /* data mode 1..13 (only 9 attaches warp fields) */
  int32_t use_antialiasing;     /* data_generation_param.use_antialiasing */
  int32_t n_tex, tex_w, tex_h;  /* texture pool: n_tex x 3 x tex_h x tex_w, planar, channel order as stored by the reference after its R<->B swap */
  int32_t n_fields;             /* mode 9: injected pool of (flow, iflow) pairs, each 2 x (H+1) x (W+1) float */
  int32_t n_threads;            /* first_level_threads */
  int32_t faithful_copies;      /* 1: also perform the reference's redundant whole-image copies (CPU-baseline timing) */
  /* pools of mixed texture sizes (TextureCollection loads whatever the list names, DataGenerator.cpp:117-149):
   * when tex_sizes is non-NULL it holds n_tex (w, h) pairs, tex_offsets the byte offset of each planar texture
   * in `textures`, and tex_w / tex_h are ignored */
  const int32_t* tex_sizes;
  const uint64_t* tex_offsets;
} oracle_config;

/* Optional debug outputs (NULL to skip). Layouts:
 *   id0/id1:  n_tasks x H x W uint32, object id of the top-most non-AA-covered object
 *   masks:    n_tasks x max_objs x 4 x H x W uint8, per top-level foreground object k (z-order):
 *             [AA frame0, AA frame1, noAA frame0, noAA frame1]
 *   frames8:  n_tasks x 2 x 3 x H x W uint8 composited frames before the float conversion
 *   flow_bw:  n_tasks x 2 x H x W float, RenderCore::computeFlowImage(objects, true) (DataGenerator.cpp:801-818): flow1
 *   occlusion: n_tasks x H x W float; NOT a reference output -- the product's own definition (include/ofdg/ofdg.h),
 *             restated here from flow0 and the two index images so that both sides can be compared */
typedef struct oracle_debug {
  uint32_t* id0;
  uint32_t* id1;
  uint8_t* masks;
  int32_t max_objs;
  uint8_t* frames8;
  float* flow_bw;
  float* occlusion;
} oracle_debug;

/* Renders every task of the batch. img0/img1: n_tasks x 3 x H x W float, flow: n_tasks x 2 x H x W.
 * Returns 0 on success, nonzero on error (message via oracle_last_error). */
int oracle_render(const oracle_config* cfg, const ofdg_task_batch* tasks, const uint8_t* textures,
                  const float* fields, float* img0, float* img1, float* flow, const oracle_debug* dbg);

/* AGG-style coverage of one closed polygon given as doubles (already in screen space);
 * writes the gray8 mask the reference's draw() would produce. aa != 0 -> gamma_none, else
 * gamma_threshold(0.5). Used by the known-answer tests. */
int oracle_raster_polygon(const double* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask);
/* Same from 24.8 fixed-point vertices (what the product's host stage emits). */
int oracle_raster_fixed(const int32_t* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask);

/* getTransformedTexture on one planar u8 image (3 x h x w) with forward matrix m[6] (AGG order). */
int oracle_transform_texture(const uint8_t* in, int32_t w, int32_t h, const double* m, uint8_t* out);

/* Texture::getRandomizedCrop on one planar u8 texture (3 x th x tw) -> 3 x out_h x out_w. */
int oracle_randomized_crop(const uint8_t* tex, int32_t tw, int32_t th, int32_t out_w, int32_t out_h,
                           float angle, float zoom, int32_t shift_x, int32_t shift_y, uint8_t* out);

/* The two 256x256 composite-mask tables [u][v] (DataGenerator.cpp:606, 626). */
void oracle_composite_luts(uint8_t* add_lut, uint8_t* sub_lut);

/* Mode-9 field producer (WarpFields::CropGenerator restated, explicitly seeded): writes n_fields
 * (flow, iflow) crops, layout n x 2 x 2 x (H+1) x (W+1) float. Values may be NaN near canvas borders. */
int oracle_generate_fields(int32_t W, int32_t H, uint32_t seed, int32_t n_fields, float* out);

const char* oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
