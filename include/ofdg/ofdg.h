/*
 * ofdg.h -- C ABI of the B200-native on-the-fly optical-flow data generator.
 *
 * This is the drop-in boundary for the reference's render path: every entry point names the
 * reference interface it stands in for (paths relative to /root/reference). Plain pointers and
 * sizes only; no C++ or torch types. Every function that can fail returns 0 on success and a
 * nonzero code otherwise, with a message available from ofdg_last_error() (thread-local).
 * No exception crosses this boundary. A generator handle is bound to one CUDA device and is not
 * thread-safe: use one handle per GPU / per producer thread.
 *
 * There is no CPU fallback: every render entry point fails with OFDG_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef OFDG_OFDG_H_
#define OFDG_OFDG_H_

#include <stdint.h>

#include "ofdg/scene.h"

#ifdef __cplusplus
extern "C" {
#endif

#define OFDG_VERSION 100

enum {
  OFDG_OK = 0,
  OFDG_ERR_ARG = 1,     /* bad argument / descriptor the reference would also reject (std::runtime_error there) */
  OFDG_ERR_CUDA = 2,    /* CUDA runtime error, or no usable device */
  OFDG_ERR_STATE = 3    /* call order (e.g. render before textures were uploaded) */
};

typedef struct ofdg_generator ofdg_generator; /* replaces DataGenerator::DataGenerator, include/caffe/data_generation/DataGenerator.h:449-500 */
typedef struct ofdg_params ofdg_params;       /* replaces ObjectParametersGenerator, DataGenerator.h:508-588 */
typedef struct ofdg_tasks ofdg_tasks;         /* an owned, growable ofdg_task_batch (std::queue<TaskBucket*> m_undone_tasks) */
typedef struct ofdg_prepared ofdg_prepared;   /* a flattened batch resident in HBM */

/* data_generation_param / data_param subset + the compile-time size of the reference
 * (src/caffe/proto/caffe.proto:6-12, DataGenerator.h:55-56). */
typedef struct ofdg_config {
  int32_t device;            /* CUDA device ordinal */
  int32_t width, height;     /* DGEN_WIDTH, DGEN_HEIGHT (512 x 384); width % 4 == 0 */
  int32_t mode;              /* 1..13 */
  int32_t use_antialiasing;  /* default 1 */
  int32_t max_batch;         /* largest batch a single render call may carry */
  int32_t reserved[8];       /* zero */
} ofdg_config;

const char* ofdg_last_error(void);
int32_t ofdg_version(void);

/* ---- host parameter stream ------------------------------------------------------------------ */
/* ObjectParametersGenerator ctor (src/caffe/DataGenerator.cpp:1358-2054): 45 engines seeded
 * seed_offset+0..44. n_fields > 0 makes mode 9 assign field-pool ids. fg_override > 0 forces
 * the number of foreground objects (stress config only; 0 = the mode's own distribution). */
int ofdg_params_create(int32_t mode, int32_t width, int32_t height, int32_t seed_offset, int32_t n_fields,
                       int32_t fg_override, ofdg_params** out);
void ofdg_params_destroy(ofdg_params* p);
/* The commission loop of load_batch (src/caffe/layers/data_generation_layer.cpp:197-214):
 * appends n_tasks freshly drawn tasks to `out`. */
int ofdg_params_generate(ofdg_params* p, int32_t n_tasks, ofdg_tasks* out);
/* Colour/noise augmentation records for subsequently generated tasks (not in the reference; specification in
 * ofdg/scene.h; drawn from five extra engines seeded 0x40000000 + seed_offset + 0..4, away from every rank's 45 reference seeds). Off by default. */
int ofdg_params_enable_augmentation(ofdg_params* p, int32_t enable);
int ofdg_params_skip(ofdg_params* p, uint64_t n_tasks);          /* checkpoint/resume: fast-forward */
/* Helper threads that produce the engines' values ahead of ofdg_params_generate's sequential walk (one job per engine and
 * batch; the k-th value of an engine depends on its seed and k only, so the stream is unchanged). 0 = off (default). */
int ofdg_params_set_threads(ofdg_params* p, int32_t threads);
uint64_t ofdg_params_tasks_generated(const ofdg_params* p);
/* Mode 9: warp-field picks made so far (pool slot of pick k = (k / 3) % n_fields). */
uint64_t ofdg_params_field_draws(const ofdg_params* p);
uint64_t ofdg_params_draws(const ofdg_params* p, int32_t slot);  /* draws made by engine `slot` so far */
const char* ofdg_params_slot_name(int32_t slot);

int ofdg_tasks_create(ofdg_tasks** out);
void ofdg_tasks_destroy(ofdg_tasks* t);
void ofdg_tasks_clear(ofdg_tasks* t);
/* Borrowed view; valid until the next mutation of `t`. */
int ofdg_tasks_view(const ofdg_tasks* t, ofdg_task_batch* out);
/* Deep-copies a caller-built batch (tests, foreign producers). */
int ofdg_tasks_assign(ofdg_tasks* t, const ofdg_task_batch* src);

/* ---- host geometry (exposed for known-answer tests) ------------------------------------------- */
/* Vertices the rasteriser is fed for an ellipse / polygon under matrix m[6] = {sx,shy,shx,sy,tx,ty}
 * (conv_transform + conv_curve + iround(v*256); DataGenerator.cpp:465-534). Writes up to `cap`
 * (x,y) int32 pairs, returns the count (or -1). */
int32_t ofdg_flatten_ellipse(double rx, double ry, const double* m, int32_t* xy, int32_t cap);
int32_t ofdg_flatten_polygon(const int32_t* seg_type, const float* seg_x, const float* seg_y, int32_t n,
                             const double* m, int32_t* xy, int32_t cap);

/* The render kernel's tile rasteriser executed on the HOST (same source, csrc/raster_tile.h) for a
 * closed outline of n 24.8 fixed-point vertices; writes the W x H gray8 mask draw() would produce
 * (aa != 0: gamma_none, else gamma_threshold(0.5)). Lets the CPU test-suite check the closed-form
 * cell arithmetic against the oracle's sequential AGG port without a GPU. */
int ofdg_debug_raster_host(const int32_t* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask);
/* Test hook (no GPU needed): the host routine that widens byte planes into float blobs for the
 * host-blob entry points, dst[i] = (float)src[i]; `streaming` selects non-temporal stores. */
int ofdg_debug_expand_host(const uint8_t* src, float* dst, uint64_t n, int32_t streaming);

/* ---- generator ------------------------------------------------------------------------------------ */
int ofdg_create(const ofdg_config* cfg, ofdg_generator** out);   /* DataGenerator ctor + Start(), DataGenerator.cpp:990-1030 */
void ofdg_destroy(ofdg_generator* g);                            /* Stop() + dtor */

/* TextureCollection ctor (DataGenerator.cpp:117-149) minus the image-file decoding: `planar` is
 * n x 3 x h x w uint8 in the channel order the reference holds after its R<->B swap. One-time
 * upload into the HBM-resident pool. ofdg_upload_textures replaces the pool; ofdg_add_textures
 * appends, so a pool may mix sizes (one call per size), like the reference's texture lists.
 * Any size from 2 x 2 works, with Texture::getRandomizedCrop's two branches (DataGenerator.cpp:87-109):
 * foreground objects use the centre W x H window of a texture that is at least W x H, else the
 * whole texture resized to W x H (resized once, here); the background crops textures that are at
 * least 2W x 2H and resizes smaller ones whole. */
int ofdg_upload_textures(ofdg_generator* g, const uint8_t* planar, int32_t n, int32_t w, int32_t h);
int ofdg_add_textures(ofdg_generator* g, const uint8_t* planar, int32_t n, int32_t w, int32_t h);
int ofdg_clear_textures(ofdg_generator* g);
/* Fills the pool with n procedural textures generated on the device (no texture database is
 * available offline); bit-identical to ofdg_b200.synth_textures() in numpy. */
int ofdg_synth_textures(ofdg_generator* g, int32_t n, int32_t w, int32_t h, uint64_t seed);
int ofdg_texture_size(const ofdg_generator* g, int32_t index, int32_t* w, int32_t* h);
int ofdg_download_texture(ofdg_generator* g, int32_t index, uint8_t* planar_out);
/* The 3 x H x W foreground view of pool texture `index` as the renderer reads it (parity checks). */
int ofdg_download_foreground_view(ofdg_generator* g, int32_t index, uint8_t* planar_out);

/* Mode 9: inject the pool of (flow, iflow) crops the reference takes from
 * WarpFields::CropGenerator::get_crop (src/caffe/WarpFields.cpp:516-538). fields is
 * n x 2 x 2 x (height+1) x (width+1) float: [field][flow|iflow][channel][y][x]. */
int ofdg_set_fields(ofdg_generator* g, const float* fields, int32_t n);

/* WarpFields::CropGenerator (src/caffe/WarpFields.cpp:469-641) on the GPU: draws the random displacer scene
 * with std::mt19937(seed) (the reference: std::random_device), builds the forward/inverse fields on the
 * 3*max(W,H) canvas with 17 self-compositions each and installs n crops as the generator's pool
 * (like ofdg_set_fields). fields_out, if not NULL, receives a host copy (same layout as ofdg_set_fields). */
int ofdg_generate_fields(ofdg_generator* g, uint32_t seed, int32_t n, float* fields_out);
/* The reference's CropGenerator keeps producing crops while training runs; every crop is handed out three times and then
 * dropped (WarpFields.cpp:516-538, 540-641). Here the pool is a ring: this call regenerates slots [first_slot, first_slot + n)
 * in place from a fresh displacer scene (std::mt19937(seed)), on a stream of its own with persistent work buffers, and updates
 * their reach. It may run on a producer thread beside render calls as long as no batch that is prepared or being rendered
 * uses those slots (the caller's ring arithmetic, csrc/host/layer.cpp). Synchronises its own stream only. */
int ofdg_refresh_fields(ofdg_generator* g, uint32_t seed, int32_t first_slot, int32_t n);
/* Grows the installed pool to `total` slots (the existing ones keep their content; the new ones are to be filled by
 * ofdg_refresh_fields before a batch uses them). Set-up time only: synchronises the device. */
int ofdg_reserve_fields(ofdg_generator* g, int32_t total);

/* Process_TaskBucket (DataGenerator.cpp:1175-1254) for a whole batch, writing straight into the
 * caller's DEVICE blobs: img0/img1 = batch x 3 x H x W, flow = batch x 2 x H x W, float32.
 * `stream` is a cudaStream_t (NULL = the generator's own stream; then the call synchronises). */
int ofdg_render(ofdg_generator* g, const ofdg_task_batch* tasks, float* d_img0, float* d_img1, float* d_flow,
                void* stream);
/* Same with HOST blobs (pinned or pageable): scene upload, kernels and the device-to-host copy of
 * the three blobs all inside the call -- what Forward_cpu delivers (data_generation_layer.cpp:266-282). */
int ofdg_render_host(ofdg_generator* g, const ofdg_task_batch* tasks, float* h_img0, float* h_img1, float* h_flow);

/* Parity instrumentation: additionally returns (to HOST buffers, any may be NULL)
 *   masks: batch x max_objs x 4 x H x W uint8 -- per top-level foreground object in z-order
 *          [AA frame0, AA frame1, noAA frame0, noAA frame1]   (MovingObjectBase::m_masks_*)
 *   id0/id1: batch x H x W uint32                              (RenderCore::index_image0/1)
 *   frames8: batch x 2 x 3 x H x W uint8                       (RenderCore::frame0/1) */
int ofdg_render_debug(ofdg_generator* g, const ofdg_task_batch* tasks, float* h_img0, float* h_img1, float* h_flow,
                      uint8_t* masks, int32_t max_objs, uint32_t* id0, uint32_t* id1, uint8_t* frames8);
/* The prepared 2W x 2H background texture of every task (Texture::getRandomizedCrop(2W,2H,...),
 * DataGenerator.cpp:1186-1192), planar batch x 3 x 2H x 2W uint8; pixels outside the region the
 * renderer needs are 0 and `need` (batch x 4: x0,y0,x1,y1) reports that region. */
int ofdg_debug_background(ofdg_generator* g, const ofdg_task_batch* tasks, uint8_t* planar_out, int32_t* need);

/* The two 256x256 composite-mask tables [u][v] exactly as the render kernel evaluates
 * MovingObjectComposite::renderMasks' float rules (DataGenerator.cpp:606, 626). */
int ofdg_debug_composite_luts(ofdg_generator* g, uint8_t* add_lut, uint8_t* sub_lut);

/* Flatten + upload once, render many times (throughput measurement with inputs resident in HBM). */
int ofdg_prepare(ofdg_generator* g, const ofdg_task_batch* tasks, ofdg_prepared** out);
void ofdg_prepared_destroy(ofdg_prepared* p);
int ofdg_render_prepared(ofdg_generator* g, const ofdg_prepared* p, float* d_img0, float* d_img1, float* d_flow,
                         void* stream);
/* The same into HOST blobs, through the pipelined host-blob path (what Forward_cpu uses). */
int ofdg_render_prepared_host(ofdg_generator* g, const ofdg_prepared* p, float* h_img0, float* h_img1, float* h_flow);

/* One-call producer used by the layer: draws `batch` tasks from `p` and renders them into device blobs. */
int ofdg_generate(ofdg_generator* g, ofdg_params* p, int32_t batch, float* d_img0, float* d_img1, float* d_flow,
                  void* stream);

/* Same, delivering into HOST blobs (what Forward_cpu hands to a CPU consumer): parameter draw, flattening,
 * upload, kernels and the device-to-host copies are pipelined chunk by chunk inside the call. */
int ofdg_generate_host(ofdg_generator* g, ofdg_params* p, int32_t batch, float* h_img0, float* h_img1, float* h_flow);

/* Production mode: scene parameters are drawn ON THE DEVICE with counter-based Philox4x32-10
 * from the same mode tables and branch structure as the host stream, flattened on the device and rendered --
 * no host work, no upload. Sample i of the stream is a pure function of (mode, seed, first_sample + i), so any
 * GPU reproduces any sample. Statistically, not bitwise, equivalent to the host stream. All 13 modes; in mode 9 the
 * warp fields come from the generator's pool (ofdg_set_fields / ofdg_generate_fields), picked by one more counter-based draw. */
int ofdg_generate_philox(ofdg_generator* g, uint64_t seed, uint64_t first_sample, int32_t batch, int32_t augment,
                         float* d_img0, float* d_img1, float* d_flow, void* stream);
/* The blueprints the device stream draws for those samples, downloaded into an ordinary task batch
 * (inspection, and rendering the very same scenes through the host path / the oracle). */
int ofdg_philox_tasks(ofdg_generator* g, uint64_t seed, uint64_t first_sample, int32_t batch, int32_t augment, ofdg_tasks* out);

/* Extra tops: the outputs RenderCore holds besides frames and forward flow
 * (index_image0/1, flow1 = computeFlowImage(objects, true), DataGenerator.cpp:740-818), plus an occlusion
 * mask derived from them. All are DEVICE float blobs for `max_batch` samples; any pointer may be NULL.
 *   flow_bw    [N][2][H][W]  flow of frame 1's pixels back to frame 0: the inverse motion of the object
 *                            index_image1 names (background: I^-1, M^-1, I), the forward field added in
 *                            mode 9 exactly as getPointFlow(inverse = true) does
 *   id0, id1   [N][1][H][W]  index images as float: 1 = background, 10 + k = k-th foreground object
 *   occlusion  [N][1][H][W]  1.0 where frame 0's pixel p is not visible in frame 1, else 0.0. Not in the
 *                            reference; defined here: t = p + flow(p) in float, q = floor(t + 0.5); p is
 *                            occluded when t is NaN, q lies outside the frame, or id1(q) != id0(p).
 * Sticky: applies to every later device-blob call (ofdg_render, ofdg_render_prepared, ofdg_generate,
 * ofdg_generate_philox) until called again; pass NULL to switch the extra tops off. */
typedef struct ofdg_extra_tops {
  float* flow_bw;
  float* id0;
  float* id1;
  float* occlusion;
} ofdg_extra_tops;
int ofdg_set_extra_tops(ofdg_generator* g, const ofdg_extra_tops* tops);

/* Number of kernel launches issued by this generator so far (bench.py's gpu_launches). */
uint64_t ofdg_launch_count(const ofdg_generator* g);
/* Device time, measured with CUDA events on the launching stream, spent in the background
 * preparation kernels and in the render kernel over the render calls made since the previous
 * call of this function (synchronises; resets the accumulation). */
int ofdg_kernel_times(ofdg_generator* g, double* prep_ms, double* render_ms, int32_t* calls);
/* Of the render time the last ofdg_kernel_times call reported: the part spent in the shade kernel (the kernel that reads the
 * textures and writes the blobs; the rest is binning and mask rasterisation). 0 with OFDG_RENDER=fused. */
double ofdg_last_shade_ms(const ofdg_generator* g);
/* Same, for the pair binning and the mask rasterisation kernels; measured only when they run in line with the rest
 * (OFDG_PIPELINE=0 OFDG_RASTER_OVERLAP=0 OFDG_BIN_OVERLAP=0: the attribution pass of bench.py), else 0. */
/* Sizes of the batch rendered last, for per-kernel roofline accounting (synchronises): (object, tile) pairs whose masks the
 * raster kernel wrote (4 KB each), prepared background pixels written and the source texels under them (0 when the scene was
 * drawn on the device). */
int ofdg_last_render_stats(ofdg_generator* g, uint64_t* pairs, uint64_t* prepared_px, uint64_t* source_px);
double ofdg_last_bin_ms(const ofdg_generator* g);
double ofdg_last_raster_ms(const ofdg_generator* g);
/* Bytes of flattened scene data the last render/prepare call copied host-to-device. */
uint64_t ofdg_last_upload_bytes(const ofdg_generator* g);
/* Bytes the last ofdg_render_host / ofdg_generate_host call copied device-to-host. The frames cross
 * PCIe as bytes and are widened into the float blobs by host threads unless a sample carries the
 * float augmentation. Environment knobs, read when the generator is created / first used:
 *   OFDG_TRANSPORT=f32     float blobs cross PCIe as they are (no host widening)
 *   OFDG_HOST_THREADS=N    worker threads of the host stages (default: half the cores, at most 16)
 *   OFDG_HOST_CHUNKS=N     pipeline chunks per call (default: 4 samples per chunk, at most 16 chunks)
 *   OFDG_TRACE_HOST=1      stage timestamps of every host-blob call on stderr */
uint64_t ofdg_last_download_bytes(const ofdg_generator* g);

#ifdef __cplusplus
}
#endif
#endif /* OFDG_OFDG_H_ */
