// C-ABI implementation (include/ofdg/ofdg.h): handle management, uploads, launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "host/expand.hpp"
#include "host/flatten.hpp"
#include "host/params.hpp"
#include "ofdg/ofdg.h"
#include "raster_tile.h"
#include "philox.cuh"
#include <atomic>
#include <set>

#include "render.cuh"
#include "warpfields.cuh"

namespace {

thread_local std::string g_error;

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
struct ArgError : std::runtime_error {
  explicit ArgError(const std::string& m) : std::runtime_error(m) {}
};
struct StateError : std::runtime_error {
  explicit StateError(const std::string& m) : std::runtime_error(m) {}
};

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
  } while (0)

template <class F>
int guarded(F&& f) {
  try {
    f();
    return OFDG_OK;
  } catch (const CudaError& e) {
    g_error = e.what();
    return OFDG_ERR_CUDA;
  } catch (const StateError& e) {
    g_error = e.what();
    return OFDG_ERR_STATE;
  } catch (const std::exception& e) {
    g_error = e.what();
    return OFDG_ERR_ARG;
  }
}

// Device buffer that only ever grows.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) CK(cudaFree(p));
    p = nullptr;
    cap = 0;
    CK(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};
struct PinnedBuf {  // pinned and mapped: kernels can read it through `dev`
  void* p = nullptr;
  void* dev = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) CK(cudaFreeHost(p));
    p = nullptr;
    dev = nullptr;
    cap = 0;
    CK(cudaHostAlloc(&p, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&dev, p, 0));
    cap = bytes;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    dev = nullptr;
    cap = 0;
  }
};

// A flattened batch resident on the device.
struct DeviceScene {
  DevBuf samples, objects, shapes, verts, deform_shape, deform_field;
  int batch = 0, n_deform = 0;
  size_t pair_bound = 0;  // upper bound of the scene's (object, tile) pairs: sizes the split render path's mask buffer
  int prep_w = 0, prep_h = 0;  // largest part of a prepared background any sample needs (pixels; 0: unknown, the whole 2W x 2H)
  uint64_t prep_px = 0, prep_src_px = 0;  // prepared pixels written / source texels under them, summed over the samples (0: unknown)
  void release() { samples.release(); objects.release(); shapes.release(); verts.release(); deform_shape.release(); deform_field.release(); }
};

}  // namespace

struct ofdg_params {
  std::unique_ptr<ofdg::ParamStream> ps;
};
struct ofdg_tasks {
  ofdg::TaskBatch tb;
};
struct ofdg_prepared {
  DeviceScene scene;
  int device = 0;
  bool augmented = false;  // some sample carries the float augmentation: its frames are not byte-valued
  // Recycling (ofdg_prepared_destroy hands the device buffers back to the generator that made them instead of freeing them:
  // cudaFree synchronises the whole device, which would serialise the consumer's GPU work on every batch).
  ofdg_generator* owner = nullptr;
  mutable cudaEvent_t last_use = nullptr;  // recorded after every render call that reads the scene
  mutable bool used = false;
};

struct ofdg_generator {
  ofdg_config cfg{};
  cudaStream_t stream = nullptr;
  // CUDA-event spans around every background-preparation / render launch (roofline timing)
  struct Span { cudaEvent_t a, b; int kind; };  // kind 0 = background preparation, 1 = render (all of it), 2 = the shade kernel alone, 3 / 4 = pair binning / mask rasterisation (in-line runs only)
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  std::vector<Span> spans;
  uint64_t timed_calls = 0;
  size_t last_upload_bytes = 0, last_download_bytes = 0;
  int scratch_batch = 0;
  // device-side (Philox) parameter stream
  // (three sets: while batch k renders, batches k+1 and k+2 are drawn and flattened on a side stream)
  DevBuf ph_slots;
  struct PhiloxSet {
    DevBuf bp, seg_type, seg_x, seg_y, obj_nbp, obj_nseg, ntop;
    DevBuf n_deform;             // mode 9: device counter of warped outlines ...
    PinnedBuf n_deform_host;     //   ... and where it lands on the host
    DeviceScene scene;
    cudaEvent_t ready = nullptr, consumed = nullptr;
    bool used = false;
  };
  static constexpr int kPhSets = 3;
  PhiloxSet ph[kPhSets];
  cudaStream_t ph_stream = nullptr;
  struct PhNext { uint64_t seed = 0, first = 0; int batch = 0, augment = 0, set = 0; };
  std::vector<PhNext> ph_queue;  // batches drawn ahead on ph_stream, oldest first (at most ph_depth)
  bool ph_dirty = false;         // drawn-ahead batches were discarded and may still be running on ph_stream
  int ph_depth = 2;              // OFDG_PHILOX_DEPTH
  int ph_batch = 0;
  DevBuf rtab_pos_x, rtab_alpha_x, rtab_pos_y, rtab_alpha_y;  // CImg linear-resize tables for every source length
  DevBuf comp_lut;  // [2][256][256] composite-mask rules (additive, subtractive), filled once by composite_lut_kernel
  // texture pool
  DevBuf pool;
  int n_tex = 0;
  size_t pool_px = 0;                    // pixels in use
  std::vector<ofdg::TexInfo> tex_info;   // host copy of the table
  DevBuf tex_info_dev;
  // mode 9 fields
  DevBuf fields, fpos_x, falpha_x, fpos_y, falpha_y, mask_raw, mask_warp, field_reach_dev;
  int n_fields = 0;
  std::vector<int> field_reach;
  // per-call scene staging + scratch
  DeviceScene scene;
  PinnedBuf staging;
  DevBuf bg, tile_hits;
  // split render path: per-tile pair ranges, the pair list, the pairs' masks, control words (csrc/render.cuh)
  DevBuf tile_range, pair_list, pair_masks, pair_ctl, bg_rows, pair_rows;
  PinnedBuf pair_overflow;  // one int the binning kernel raises if a batch ever had more pairs than the host-computed bound
  int pair_cap = 0;
  bool split_render = true;  // OFDG_RENDER=fused selects the single-kernel path
  int pair_cap_limit = 0;    // OFDG_TEST_PAIR_CAP: pretend the pair buffers are this small (tests of the overflow trap)
  int philox_fg_override = 0;  // OFDG_TEST_PHILOX_FG: forced object count of the device stream (tests of its truncation flag)
  DevBuf out0, out1, outf;  // device blobs for the *_host entry points
  ofdg_extra_tops extra{};  // extra tops of the device-blob calls (ofdg_set_extra_tops)
  DevBuf ids8;              // object ranks per pixel, scratch of the occlusion pass
  // uint8 transport of the host-blob path: byte frames on the device, their pinned landing area, the
  // host threads that widen them into the caller's float blobs (host/expand.hpp), one event per chunk
  static constexpr int kMaxChunks = 32;
  DevBuf out8;
  PinnedBuf host8;
  std::unique_ptr<ofdg::HostPool> workers;
  cudaEvent_t chunk_copied[kMaxChunks] = {};
  // host stages of the pipeline: per-chunk task batches (parameter-stream flavour), one flattened part per sample
  ofdg::TaskBatch chunk_tasks[kMaxChunks];
  std::vector<ofdg::FlatBatch> sample_flat;
  struct {
    std::mutex mu;
    std::condition_variable cv;
    int done[kMaxChunks];
    std::string error;
  } host_sync;
  bool transport_u8 = true;
  DevBuf dbg_masks, dbg_id0, dbg_id1, dbg_frames8, dbg_planar;
  ofdg::FlatBatch flat;
  ofdg::TaskBatch gen_tasks;
  // host-blob pipeline: two scene/staging sets, a copy stream, events
  DeviceScene pipe_scene[2];
  PinnedBuf pipe_staging[2];
  ofdg::FlatBatch pipe_flat[2];
  cudaStream_t copy_stream = nullptr;
  cudaStream_t bin_stream = nullptr;  // pair binning of a batch beside its background preparation (OFDG_BIN_OVERLAP=0: same stream)
  // Cross-batch software pipeline (OFDG_PIPELINE=0: off): a second scratch set and a preparation stream, so that the background
  // preparation and the mask rasterisation of batch k+1 run beside the shade kernel of batch k. Set 0 is the members above
  // (bg, tile_range, pair_list, pair_masks, pair_ctl), set 1 is `alt`. Every render call records set_shade_done for the set it used.
  struct { DevBuf bg, tile_range, pair_list, pair_masks, pair_ctl, bg_rows, pair_rows; } alt;
  cudaStream_t prep_stream = nullptr;
  cudaEvent_t set_prep_done[2] = {nullptr, nullptr}, set_raster_done[2] = {nullptr, nullptr}, set_shade_done[2] = {nullptr, nullptr};
  bool set_used[2] = {false, false};
  bool pipeline = false, philox_pipeline = false;
  uint64_t pipe_calls = 0;
  cudaEvent_t bin_fork = nullptr, bin_join = nullptr;
  bool raster_overlap = false, philox_raster_overlap = false;
  cudaEvent_t pipe_uploaded[2] = {nullptr, nullptr}, pipe_rendered[2] = {nullptr, nullptr}, render_done[2] = {nullptr, nullptr};
  bool render_set_used[2] = {false, false};
  uint64_t render_calls = 0;
  std::atomic<uint64_t> launches{0};
  float last_kernel_ms = 0.f;
  double last_shade_ms = 0.0, last_bin_ms = 0.0, last_raster_ms = 0.0;  // of the spans ofdg_kernel_times summed last
  uint64_t last_prep_px = 0, last_prep_src_px = 0;  // of the scene rendered last (ofdg_last_render_stats)
  int last_set = 0;

  // ofdg_prepare: its own flatten pool, staging and upload stream, so that a producer thread can prepare the next batch
  // while another thread renders (the two only meet in the free list of recycled scenes)
  std::mutex prepare_mu, free_mu;
  std::condition_variable prepare_cv;
  std::unique_ptr<ofdg::HostPool> prep_workers;
  struct PrepCtx {  // what one ofdg_prepare call works with; a few of them, so that calls of different threads overlap
    std::vector<ofdg::FlatBatch> flat;
    PinnedBuf staging;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::condition_variable cv;
    int remaining = 0;
    std::string error;
    bool busy = false;
  };
  static constexpr int kPrepCtx = 3;
  PrepCtx prep_ctx[kPrepCtx];
  std::vector<ofdg_prepared*> free_scenes;
  static constexpr size_t kMaxFreeScenes = 12;
  // mode 9: refreshing slots of the field pool while other slots are being rendered (ofdg_refresh_fields)
  ofdg::WfScratch wf_scratch;
  cudaStream_t field_stream = nullptr;
  PinnedBuf reach_host;

  void use() const { CK(cudaSetDevice(cfg.device)); }
};

namespace {
// Generators alive in this process: a prepared scene that outlives its generator is freed instead of recycled.
std::mutex g_live_mu;
std::set<ofdg_generator*> g_live;
}  // namespace

namespace {

void check_batch(const ofdg_generator* g, int n) {
  if (n <= 0) throw ArgError("empty task batch");
  if (n > g->cfg.max_batch) throw ArgError("batch larger than ofdg_config.max_batch");
  if (g->n_tex <= 0) throw StateError("no textures uploaded (ofdg_upload_textures / ofdg_synth_textures)");
}

// Concatenates flattened parts (each with indices relative to itself) into the pinned staging area, rebasing
// the indices, and starts the host-to-device copies on stream s.
void upload_scene_parts(ofdg_generator* g, const ofdg::FlatBatch* parts, int n_parts, DeviceScene& ds, PinnedBuf& staging, cudaStream_t s) {
  size_t ns = 0, no = 0, nsh = 0, nv = 0, nd = 0;
  for (int i = 0; i < n_parts; ++i) {
    ns += parts[i].samples.size(); no += parts[i].objects.size(); nsh += parts[i].shapes.size();
    nv += parts[i].verts.size(); nd += parts[i].deform_shape.size();
  }
  const size_t b0 = ns * sizeof(ofdg::FlatSample), b1 = no * sizeof(ofdg::FlatObject), b2 = nsh * sizeof(ofdg::FlatShape),
               b3 = nv * sizeof(ofdg::FlatVertex), bd = nd * sizeof(int32_t);
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o0 = 0, o1 = al(b0), o2 = o1 + al(b1), o3 = o2 + al(b2), o4 = o3 + al(b3), o5 = o4 + al(bd);
  staging.reserve(o5 + al(bd) + 256);
  char* st = (char*)staging.p;
  ofdg::FlatSample* S = (ofdg::FlatSample*)(st + o0);
  ofdg::FlatObject* O = (ofdg::FlatObject*)(st + o1);
  ofdg::FlatShape* Sh = (ofdg::FlatShape*)(st + o2);
  ofdg::FlatVertex* V = (ofdg::FlatVertex*)(st + o3);
  int32_t* Ds = (int32_t*)(st + o4);
  int32_t* Df = (int32_t*)(st + o5);
  size_t is = 0, io = 0, ish = 0, iv = 0, id = 0;
  for (int i = 0; i < n_parts; ++i) {
    const ofdg::FlatBatch& fb = parts[i];
    if (!fb.samples.empty()) std::memcpy(S + is, fb.samples.data(), fb.samples.size() * sizeof(ofdg::FlatSample));
    if (!fb.objects.empty()) std::memcpy(O + io, fb.objects.data(), fb.objects.size() * sizeof(ofdg::FlatObject));
    if (!fb.shapes.empty()) std::memcpy(Sh + ish, fb.shapes.data(), fb.shapes.size() * sizeof(ofdg::FlatShape));
    if (!fb.verts.empty()) std::memcpy(V + iv, fb.verts.data(), fb.verts.size() * sizeof(ofdg::FlatVertex));
    if (i > 0) {
      for (size_t k = 0; k < fb.samples.size(); ++k) S[is + k].obj_begin += (int32_t)io;
      for (size_t k = 0; k < fb.objects.size(); ++k) O[io + k].shape_begin += (int32_t)ish;
      for (size_t k = 0; k < fb.shapes.size(); ++k) {
        ofdg::FlatShape& sh = Sh[ish + k];
        sh.vbegin[0] += (int32_t)iv; sh.vbegin[1] += (int32_t)iv;
        if (sh.deform >= 0) sh.deform += (int32_t)id;
      }
    }
    for (size_t k = 0; k < fb.deform_shape.size(); ++k) {
      Ds[id + k] = fb.deform_shape[k] + (int32_t)ish;
      Df[id + k] = fb.deform_field[k];
    }
    is += fb.samples.size(); io += fb.objects.size(); ish += fb.shapes.size(); iv += fb.verts.size(); id += fb.deform_shape.size();
  }
  ds.samples.reserve(b0 + 256); ds.objects.reserve(b1 + 256); ds.shapes.reserve(b2 + 256); ds.verts.reserve(b3 + 256);
  if (nd) { ds.deform_shape.reserve(bd + 256); ds.deform_field.reserve(bd + 256); }  // mode 9 only
  ds.batch = (int)ns;
  ds.n_deform = (int)nd;
  {  // (object, tile) pairs: the tiles an object's box overlaps in frame 0 or in frame 1 -- two rectangles of tiles, counted as
     // |A| + |B| - |A and B|, which is exactly what bin_pairs_kernel's box tests count (the pair buffers are sized from it)
    const int tiles_x = (g->cfg.width + ofdg::TW - 1) / ofdg::TW, tiles_y = (g->cfg.height + ofdg::TH - 1) / ofdg::TH;
    size_t pairs = 0;
    for (int i = 0; i < n_parts; ++i)
      for (const ofdg::FlatObject& o : parts[i].objects) {
        int c0[2], c1[2], r0[2], r1[2];
        size_t n[2];
        for (int f = 0; f < 2; ++f) {
          c0[f] = std::max(0, o.bbox[f][0] >= 0 ? o.bbox[f][0] / ofdg::TW : 0); c1[f] = std::min(tiles_x - 1, o.bbox[f][2] >= 0 ? o.bbox[f][2] / ofdg::TW : -1);
          r0[f] = std::max(0, o.bbox[f][1] >= 0 ? o.bbox[f][1] / ofdg::TH : 0); r1[f] = std::min(tiles_y - 1, o.bbox[f][3] >= 0 ? o.bbox[f][3] / ofdg::TH : -1);
          n[f] = (c1[f] >= c0[f] && r1[f] >= r0[f]) ? (size_t)(c1[f] - c0[f] + 1) * (r1[f] - r0[f] + 1) : 0;
        }
        size_t both = 0;
        if (n[0] && n[1]) {
          const int ca = std::max(c0[0], c0[1]), cb = std::min(c1[0], c1[1]), ra = std::max(r0[0], r0[1]), rb = std::min(r1[0], r1[1]);
          if (cb >= ca && rb >= ra) both = (size_t)(cb - ca + 1) * (rb - ra + 1);
        }
        pairs += n[0] + n[1] - both;
      }
    ds.pair_bound = pairs;
  }
  ds.prep_w = ds.prep_h = 0;  // the preparation kernel's grid covers the largest needed region, not the whole 2W x 2H canvas
  ds.prep_px = ds.prep_src_px = 0;
  for (int i = 0; i < n_parts; ++i)
    for (const ofdg::FlatSample& fs : parts[i].samples) {
      const int nw = fs.prep.need[2] - fs.prep.need[0] + 1, nh = fs.prep.need[3] - fs.prep.need[1] + 1;
      ds.prep_w = std::max(ds.prep_w, nw);
      ds.prep_h = std::max(ds.prep_h, nh);
      if (nw > 0 && nh > 0) {
        ds.prep_px += (uint64_t)nw * nh;
        ds.prep_src_px += (uint64_t)((double)nw * nh * ((double)fs.prep.crop_w / (2.0 * g->cfg.width)) * ((double)fs.prep.crop_h / (2.0 * g->cfg.height)));
      }
    }
  ofdg::UploadSegments u{};
  const char* dv = (const char*)staging.dev;
  auto seg = [&u, dv](void* dst, size_t off, size_t bytes) {
    if (!bytes) return;
    u.src[u.n] = dv + off; u.dst[u.n] = dst; u.n16[u.n] = (unsigned)((bytes + 15) / 16);
    ++u.n;
  };
  seg(ds.samples.p, o0, b0); seg(ds.objects.p, o1, b1); seg(ds.shapes.p, o2, b2); seg(ds.verts.p, o3, b3);
  if (nd) { seg(ds.deform_shape.p, o4, bd); seg(ds.deform_field.p, o5, bd); }
  g->launches += ofdg::launch_scene_upload(u, s);
  g->last_upload_bytes = b0 + b1 + b2 + b3 + 2 * bd;
}

void upload_scene(ofdg_generator* g, const ofdg::FlatBatch& fb, DeviceScene& ds, PinnedBuf& staging, cudaStream_t s) {
  upload_scene_parts(g, &fb, 1, ds, staging, s);
}

ofdg::FlattenConfig flatten_config(const ofdg_generator* g) {
  ofdg::FlattenConfig fc;
  fc.W = g->cfg.width; fc.H = g->cfg.height;
  fc.tex_info = g->tex_info.data(); fc.n_tex = g->n_tex;
  fc.mode = g->cfg.mode;
  fc.n_fields = g->n_fields;
  fc.field_reach = g->field_reach.empty() ? nullptr : g->field_reach.data();
  return fc;
}

void flatten_tasks(ofdg_generator* g, const ofdg_task_batch* tasks, ofdg::FlatBatch* into = nullptr) {
  ofdg::FlatBatch& flat = into ? *into : g->flat;
  flat.clear();
  ofdg::flatten(*tasks, flatten_config(g), flat);
}

void ensure_scratch(ofdg_generator* g, int batch) {
  if (batch <= g->scratch_batch) return;
  CK(cudaDeviceSynchronize());  // growing the scratch while earlier launches may still read it: drain first
  const size_t W = g->cfg.width, H = g->cfg.height;
  g->bg.reserve((size_t)batch * 4 * W * H * sizeof(uchar4));
  g->tile_hits.reserve(ofdg::tile_hits_bytes(batch, (int)W, (int)H));
  if (g->split_render) {
    const size_t tiles = ofdg::tile_hits_bytes(1, (int)W, (int)H) / ofdg::TILE_HIT_STRIDE;
    g->tile_range.reserve((size_t)batch * tiles * sizeof(int2));
    g->pair_ctl.reserve(4 * sizeof(int));
    g->bg_rows.reserve((size_t)batch * H * 2 * sizeof(int4));
    if (g->pipeline) {
      g->alt.bg_rows.reserve((size_t)batch * H * 2 * sizeof(int4));
      g->alt.bg.reserve((size_t)batch * 4 * W * H * sizeof(uchar4));
      g->alt.tile_range.reserve((size_t)batch * tiles * sizeof(int2));
      g->alt.pair_ctl.reserve(4 * sizeof(int));
    }
    if (!g->pair_overflow.p) {
      g->pair_overflow.reserve(2 * sizeof(int));  // [0] pair buffer overflow, [1] a device-drawn scene did not fit its fixed strides
      ((volatile int*)g->pair_overflow.p)[0] = 0;
      ((volatile int*)g->pair_overflow.p)[1] = 0;
    }
  }
  g->scratch_batch = batch;
}

ofdg::RenderArgs make_args(ofdg_generator* g, const DeviceScene& ds, float* d0, float* d1, float* df, int set = 0) {
  ofdg::RenderArgs a{};
  g->last_prep_px = ds.prep_px; g->last_prep_src_px = ds.prep_src_px; g->last_set = set;
  a.samples = (const ofdg::FlatSample*)ds.samples.p;
  a.objects = (const ofdg::FlatObject*)ds.objects.p;
  a.shapes = (const ofdg::FlatShape*)ds.shapes.p;
  a.verts = (const ofdg::FlatVertex*)ds.verts.p;
  a.batch = ds.batch;
  a.W = g->cfg.width; a.H = g->cfg.height;
  a.prep_w = ds.prep_w; a.prep_h = ds.prep_h;
  a.float_bias = 0x4B000000u;
  a.use_aa = g->cfg.use_antialiasing;
  a.pool = (const uchar4*)g->pool.p;
  a.tex_info = (const ofdg::TexInfo*)g->tex_info_dev.p;
  a.bg = (uchar4*)(set ? g->alt.bg.p : g->bg.p);
  a.tile_hits = (uint8_t*)g->tile_hits.p;
  if (g->split_render) {
    if (ds.pair_bound > (size_t)g->pair_cap) {  // grow the pair buffers (4 KB of masks per pair); earlier launches may still use the old ones
      CK(cudaDeviceSynchronize());
      const size_t cap = std::max<size_t>(ds.pair_bound + ds.pair_bound / 4, 4096);
      g->pair_list.reserve(cap * sizeof(int4));
      g->pair_masks.reserve(cap * ofdg::pair_mask_bytes_per_pair());
      g->pair_rows.reserve(cap * ofdg::pair_row_bytes_per_pair());
      if (g->pipeline) {
        g->alt.pair_rows.reserve(cap * ofdg::pair_row_bytes_per_pair());
        g->alt.pair_list.reserve(cap * sizeof(int4));
        g->alt.pair_masks.reserve(cap * ofdg::pair_mask_bytes_per_pair());
      }
      g->pair_cap = (int)cap;
    }
    if (set) {
      a.tile_range = (int2*)g->alt.tile_range.p; a.pair_list = (int4*)g->alt.pair_list.p; a.pair_masks = (uint32_t*)g->alt.pair_masks.p;
      a.pair_ctl = (int*)g->alt.pair_ctl.p; a.bg_rows = (int4*)g->alt.bg_rows.p; a.pair_rows = (int4*)g->alt.pair_rows.p;
    } else {
      a.tile_range = (int2*)g->tile_range.p; a.pair_list = (int4*)g->pair_list.p; a.pair_masks = (uint32_t*)g->pair_masks.p;
      a.pair_ctl = (int*)g->pair_ctl.p; a.bg_rows = (int4*)g->bg_rows.p; a.pair_rows = (int4*)g->pair_rows.p;
    }
    a.pair_cap = g->pair_cap_limit > 0 ? std::min(g->pair_cap, g->pair_cap_limit) : g->pair_cap;
    a.pair_overflow = (int*)g->pair_overflow.dev;
    a.comp_lut = (const uint8_t*)g->comp_lut.p;
  }
  a.pos_x = (const int*)g->rtab_pos_x.p; a.alpha_x = (const double*)g->rtab_alpha_x.p;
  a.pos_y = (const int*)g->rtab_pos_y.p; a.alpha_y = (const double*)g->rtab_alpha_y.p;
  a.fields = (const float*)g->fields.p;
  a.n_fields = g->n_fields;
  a.n_deform = ds.n_deform;
  if (ds.n_deform) {
    const size_t P = (size_t)g->cfg.width * g->cfg.height;
    g->mask_raw.reserve((size_t)ds.n_deform * 2 * P);
    g->mask_warp.reserve((size_t)ds.n_deform * 2 * P);
  }
  a.deform_shape = (const int*)ds.deform_shape.p;
  a.deform_field = (const int*)ds.deform_field.p;
  a.mask_raw = (uint8_t*)g->mask_raw.p;
  a.mask_warp = (uint8_t*)g->mask_warp.p;
  a.fpos_x = (const int*)g->fpos_x.p; a.falpha_x = (const double*)g->falpha_x.p;
  a.fpos_y = (const int*)g->fpos_y.p; a.falpha_y = (const double*)g->falpha_y.p;
  a.img0 = d0; a.img1 = d1; a.flow = df;
  return a;
}

// Device-blob calls also fill the extra tops the caller registered (ofdg_set_extra_tops).
ofdg::RenderArgs with_extra_tops(ofdg_generator* g, ofdg::RenderArgs a) {
  a.flow_bw = g->extra.flow_bw; a.top_id0 = g->extra.id0; a.top_id1 = g->extra.id1; a.occlusion = g->extra.occlusion;
  if (a.occlusion) {
    g->ids8.reserve((size_t)g->cfg.max_batch * 2 * g->cfg.width * g->cfg.height);
    a.ids8 = (uint8_t*)g->ids8.p;
  }
  return a;
}

// After a synchronisation: the pair buffers are sized from an upper bound of the batch's (object, tile) pairs, so the
// binning kernel can never run out of room; if it ever did (its kernels then skip the batch instead of writing out of
// bounds) the call that notices fails loudly rather than handing back blobs that were not rendered.
void check_pair_overflow(ofdg_generator* g) {
  volatile int* f = (volatile int*)g->pair_overflow.p;
  if (f && f[1]) {
    f[1] = 0;
    throw StateError("device parameter stream: a sample asked for more than 32 foreground objects or an outline for more than 512 "
                     "vertices per frame; it was rendered without them (use the host parameter stream for such scenes)");
  }
  if (f && f[0]) {
    f[0] = 0;
    throw StateError("internal: more (object, tile) pairs than the mask buffer was sized for; the batch was not rendered");
  }
}

cudaEvent_t timing_event(ofdg_generator* g) {
  if (g->ev_next == g->ev_pool.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    g->ev_pool.push_back(e);
  }
  return g->ev_pool[g->ev_next++];
}

// Background preparation + render of one batch on stream s, bracketed by events. The pair binning and (side_raster) the mask
// rasterisation are forked onto the high-priority side stream beside the preparation and joined before the shade kernel.
// (Running the preparation of the NEXT chunk on a second stream next to the render kernels was measured and is slower: the
// extra launches cost more than the overlap gains -- profiles/README.md.)
// Pipelined form (pipe_set >= 0; the caller built `a` with make_args(.., pipe_set)): the scene is already resident (scene_ready, if
// given, says when), so nothing of this batch's front end depends on what is queued on s. The pair binning + mask rasterisation
// go to the high-priority side stream and the background preparation to the preparation stream, both gated only by the shade
// kernel that last read this scratch set (two calls ago); s waits for both and runs the shade kernel. Queued back to back, the
// front end of batch k+1 therefore runs beside the shade kernel of batch k.
void run_kernels_pipelined(ofdg_generator* g, const ofdg::RenderArgs& a, cudaStream_t s, int set, cudaEvent_t scene_ready) {
  if (g->spans.size() > 60000) { g->spans.clear(); g->ev_next = 0; g->timed_calls = 0; }
  cudaStream_t front[2] = {g->bin_stream, g->prep_stream};
  for (cudaStream_t f : front) {
    if (scene_ready) CK(cudaStreamWaitEvent(f, scene_ready, 0));
    if (g->set_used[set]) CK(cudaStreamWaitEvent(f, g->set_shade_done[set], 0));
  }
  g->launches += ofdg::launch_bin_pairs(a, g->bin_stream);
  g->launches += ofdg::launch_raster_pairs(a, g->bin_stream);
  CK(cudaEventRecord(g->set_raster_done[set], g->bin_stream));
  ofdg_generator::Span sp{timing_event(g), timing_event(g), 0};
  CK(cudaEventRecord(sp.a, g->prep_stream));
  g->launches += ofdg::launch_background_prep(a, g->prep_stream);
  CK(cudaEventRecord(sp.b, g->prep_stream));
  CK(cudaEventRecord(g->set_prep_done[set], g->prep_stream));
  g->spans.push_back(sp);
  CK(cudaStreamWaitEvent(s, g->set_raster_done[set], 0));
  CK(cudaStreamWaitEvent(s, g->set_prep_done[set], 0));
  ofdg_generator::Span sr{timing_event(g), timing_event(g), 1};
  CK(cudaEventRecord(sr.a, s));  // completes once both waits are satisfied: the span is the shade kernel (+ the occlusion pass)
  g->launches += ofdg::launch_render_split(a, s, nullptr, 2);
  CK(cudaEventRecord(sr.b, s));
  g->spans.push_back(ofdg_generator::Span{sr.a, sr.b, 2});
  g->spans.push_back(sr);
  CK(cudaEventRecord(g->set_shade_done[set], s));
  g->set_used[set] = true;
  ++g->timed_calls;
  CK(cudaGetLastError());
}

// Whether a render call can take the pipelined form, and with which scratch set (-1: no). Warped outlines (mode 9) keep the
// in-order form: their pre-pass scratch is single-buffered.
int pipeline_set(ofdg_generator* g, const DeviceScene& ds) {
  if (!g->pipeline || ds.n_deform) return -1;
  return (int)(g->pipe_calls++ & 1);
}

void run_kernels(ofdg_generator* g, const ofdg::RenderArgs& a, cudaStream_t s, bool deform_prepass = true, bool side_raster = true) {
  if (g->spans.size() > 60000) { g->spans.clear(); g->ev_next = 0; g->timed_calls = 0; }  // nobody is reading the timings
  if (g->pipeline && g->set_used[0]) CK(cudaStreamWaitEvent(s, g->set_shade_done[0], 0));  // in-order calls use scratch set 0
  if (deform_prepass) g->launches += ofdg::launch_deform_prepass(a, s);
  ofdg_generator::Span sp{timing_event(g), timing_event(g), 0};
  CK(cudaEventRecord(sp.a, s));
  const bool fork_bin = g->bin_stream && a.pair_ctl && a.flow;
  const bool fork_raster = fork_bin && g->raster_overlap && side_raster;
  if (fork_bin) {  // everything queued on s so far (scene upload, the previous batch's kernels) precedes the binning
    CK(cudaEventRecord(g->bin_fork, s));
    CK(cudaStreamWaitEvent(g->bin_stream, g->bin_fork, 0));
    g->launches += ofdg::launch_bin_pairs(a, g->bin_stream);  // queued ahead of the preparation's blocks: runs beside them
    if (fork_raster) g->launches += ofdg::launch_raster_pairs(a, g->bin_stream);  // masks need the scene + pair list only
    CK(cudaEventRecord(g->bin_join, g->bin_stream));
  }
  if (!a.pair_ctl || !a.flow) g->launches += ofdg::launch_bin(a, s);  // (the split path bins inside launch_render_split)
  g->launches += ofdg::launch_background_prep(a, s);
  CK(cudaEventRecord(sp.b, s));
  g->spans.push_back(sp);
  if (a.flow) {
    ofdg_generator::Span sr{timing_event(g), timing_event(g), 1};
    CK(cudaEventRecord(sr.a, s));
    if (a.pair_ctl) {
      cudaEvent_t mid = timing_event(g);
      if (fork_bin) CK(cudaStreamWaitEvent(s, g->bin_join, 0));
      int done = fork_raster ? 2 : fork_bin ? 1 : 0;
      if (done == 0) {  // everything in line (OFDG_BIN_OVERLAP=0): the binning and the rasterisation get spans of their own
        cudaEvent_t e1 = timing_event(g);
        g->launches += ofdg::launch_bin_pairs(a, s);
        CK(cudaEventRecord(e1, s));
        g->launches += ofdg::launch_raster_pairs(a, s);
        g->spans.push_back(ofdg_generator::Span{sr.a, e1, 3});
        g->spans.push_back(ofdg_generator::Span{e1, mid, 4});
        done = 2;
      }
      g->launches += ofdg::launch_render_split(a, s, mid, done);
      CK(cudaEventRecord(sr.b, s));
      g->spans.push_back(ofdg_generator::Span{mid, sr.b, 2});
    } else {
      g->launches += ofdg::launch_render(a, s);
      CK(cudaEventRecord(sr.b, s));
    }
    g->spans.push_back(sr);
  }
  if (g->pipeline) {  // a later pipelined call that picks set 0 must not start its front end before this call is done with it
    CK(cudaEventRecord(g->set_shade_done[0], s));
    g->set_used[0] = true;
  }
  ++g->timed_calls;
  CK(cudaGetLastError());
}

}  // namespace

extern "C" {

const char* ofdg_last_error(void) { return g_error.c_str(); }
int32_t ofdg_version(void) { return OFDG_VERSION; }

// ---- parameter stream ---------------------------------------------------------------------------
int ofdg_params_create(int32_t mode, int32_t width, int32_t height, int32_t seed_offset, int32_t n_fields,
                       int32_t fg_override, ofdg_params** out) {
  return guarded([&] {
    if (!out) throw ArgError("null output pointer");
    if (width <= 0 || height <= 0) throw ArgError("bad output size");
    std::unique_ptr<ofdg_params> p(new ofdg_params);
    p->ps.reset(new ofdg::ParamStream(mode, width, height, seed_offset, n_fields, fg_override));
    *out = p.release();
  });
}
void ofdg_params_destroy(ofdg_params* p) { delete p; }
int ofdg_params_generate(ofdg_params* p, int32_t n_tasks, ofdg_tasks* out) {
  return guarded([&] {
    if (!p || !out || n_tasks < 0) throw ArgError("bad arguments");
    p->ps->next_tasks(out->tb, n_tasks);
  });
}
int ofdg_params_set_threads(ofdg_params* p, int32_t threads) {
  return guarded([&] {
    if (!p || threads < 0) throw ArgError("bad arguments");
    p->ps->set_lookahead_threads(threads);
  });
}
int ofdg_params_skip(ofdg_params* p, uint64_t n_tasks) {
  return guarded([&] {
    if (!p) throw ArgError("null stream");
    p->ps->skip(n_tasks);
  });
}
int ofdg_params_enable_augmentation(ofdg_params* p, int32_t enable) {
  return guarded([&] {
    if (!p) throw ArgError("null stream");
    p->ps->enable_augmentation(enable != 0);
  });
}
uint64_t ofdg_params_tasks_generated(const ofdg_params* p) { return p ? p->ps->tasks_generated() : 0; }
uint64_t ofdg_params_field_draws(const ofdg_params* p) { return p ? p->ps->field_draws() : 0; }
uint64_t ofdg_params_draws(const ofdg_params* p, int32_t slot) {
  return (p && slot >= 0 && slot < ofdg::kNumSlots) ? p->ps->draws(slot) : 0;
}
const char* ofdg_params_slot_name(int32_t slot) { return ofdg::slot_name(slot); }

int ofdg_tasks_create(ofdg_tasks** out) {
  return guarded([&] {
    if (!out) throw ArgError("null output pointer");
    *out = new ofdg_tasks;
  });
}
void ofdg_tasks_destroy(ofdg_tasks* t) { delete t; }
void ofdg_tasks_clear(ofdg_tasks* t) {
  if (t) t->tb.clear();
}
int ofdg_tasks_view(const ofdg_tasks* t, ofdg_task_batch* out) {
  return guarded([&] {
    if (!t || !out) throw ArgError("null pointer");
    *out = t->tb.view();
  });
}
int ofdg_tasks_assign(ofdg_tasks* t, const ofdg_task_batch* s) {
  return guarded([&] {
    if (!t || !s) throw ArgError("null pointer");
    t->tb.task_begin.assign(s->task_begin, s->task_begin + s->n_tasks + 1);
    t->tb.blueprints.assign(s->blueprints, s->blueprints + s->n_blueprints);
    t->tb.seg_type.assign(s->seg_type, s->seg_type + s->n_segments);
    t->tb.seg_x.assign(s->seg_x, s->seg_x + s->n_segments);
    t->tb.seg_y.assign(s->seg_y, s->seg_y + s->n_segments);
    if (s->augment) t->tb.augment.assign(s->augment, s->augment + s->n_tasks);
    else t->tb.augment.clear();
  });
}

// ---- host geometry ---------------------------------------------------------------------------------
int32_t ofdg_flatten_ellipse(double rx, double ry, const double* m, int32_t* xy, int32_t cap) {
  try {
    std::vector<ofdg::FlatVertex> v;
    ofdg::flatten_ellipse(rx, ry, m, v);
    const int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; }
    return n;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}
int32_t ofdg_flatten_polygon(const int32_t* seg_type, const float* seg_x, const float* seg_y, int32_t n,
                             const double* m, int32_t* xy, int32_t cap) {
  try {
    std::vector<ofdg::FlatVertex> v;
    ofdg::flatten_polygon(seg_type, seg_x, seg_y, n, m, v);
    const int k = (int)v.size();
    for (int i = 0; i < k && i < cap; ++i) { xy[2 * i] = v[i].x; xy[2 * i + 1] = v[i].y; }
    return k;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// The tile rasteriser of the render kernel (raster_tile.h) run on the host over every tile of a
// W x H frame: same code path as the device, atomics replaced by plain adds.
int ofdg_debug_raster_host(const int32_t* xy, int32_t n, int32_t W, int32_t H, int32_t aa, uint8_t* mask) {
  return guarded([&] {
    if (!xy || !mask || n < 1 || W <= 0 || H <= 0) throw ArgError("bad arguments");
    using namespace ofdg;
    std::vector<int> cover(TH * TW), area(TH * TW), carry(TH);
    for (int ty0 = 0; ty0 < H; ty0 += TH)
      for (int tx0 = 0; tx0 < W; tx0 += TW) {
        std::fill(cover.begin(), cover.end(), 0);
        std::fill(area.begin(), area.end(), 0);
        std::fill(carry.begin(), carry.end(), 0);
        for (int e = 0; e < n; ++e) {
          const int e2 = e + 1 == n ? 0 : e + 1;
          tile_edge<false>(cover.data(), area.data(), carry.data(), tx0, ty0, xy[2 * e], xy[2 * e + 1], xy[2 * e2], xy[2 * e2 + 1]);
        }
        for (int r = 0; r < TH && ty0 + r < H; ++r) {
          int cum = carry[r];
          for (int c = 0; c < TW && tx0 + c < W; ++c) {
            cum += cover[r * TW + c];
            const int cv = coverage_alpha(cum, area[r * TW + c]);
            mask[(size_t)(ty0 + r) * W + tx0 + c] = (uint8_t)(aa ? graylut((unsigned)cv) : (cv >= 128 ? 255u : 0u));
          }
        }
      }
  });
}

int ofdg_debug_expand_host(const uint8_t* src, float* dst, uint64_t n, int32_t streaming) {
  return guarded([&] {
    if ((!src || !dst) && n) throw ArgError("null pointer");
    ofdg::expand_u8_to_f32(src, dst, (size_t)n, streaming != 0);
  });
}

// ---- generator ---------------------------------------------------------------------------------------
int ofdg_create(const ofdg_config* cfg, ofdg_generator** out) {
  return guarded([&] {
    if (!cfg || !out) throw ArgError("null pointer");
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->width % 4) throw ArgError("width/height must be positive and width a multiple of 4");
    if (cfg->mode < 1 || cfg->mode > 13) throw ArgError("BAD MODE");
    if (cfg->max_batch <= 0 || cfg->max_batch > 65535) throw ArgError("max_batch must be in 1..65535");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw CudaError(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) throw ArgError("bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) throw CudaError("this library carries sm_100a code only; found sm_" + std::to_string(prop.major * 10 + prop.minor));
    std::unique_ptr<ofdg_generator> g(new ofdg_generator);
    g->cfg = *cfg;
    CK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    {
      const char* pp = std::getenv("OFDG_PHILOX_PRIORITY");  // the look-ahead parameter kernels are placed ahead of the render's (0: not)
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&g->ph_stream, cudaStreamNonBlocking, (pp && std::string(pp) == "0") ? 0 : hi));
    }
    if (const char* t = std::getenv("OFDG_PHILOX_DEPTH")) g->ph_depth = std::max(1, std::min(std::atoi(t), ofdg_generator::kPhSets - 1));
    for (int i = 0; i < ofdg_generator::kPhSets; ++i) {
      CK(cudaEventCreateWithFlags(&g->ph[i].ready, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->ph[i].consumed, cudaEventDisableTiming));
    }
    {  // one-time: the resize tables for every possible crop length (they depend on nothing else)
      const size_t W2 = 2 * (size_t)cfg->width, H2 = 2 * (size_t)cfg->height;
      g->rtab_pos_x.reserve(W2 * W2 * sizeof(int)); g->rtab_alpha_x.reserve(W2 * W2 * sizeof(double));
      g->rtab_pos_y.reserve(H2 * H2 * sizeof(int)); g->rtab_alpha_y.reserve(H2 * H2 * sizeof(double));
      ofdg::launch_resize_tables((int*)g->rtab_pos_x.p, (double*)g->rtab_alpha_x.p, (int)W2, g->stream);
      ofdg::launch_resize_tables((int*)g->rtab_pos_y.p, (double*)g->rtab_alpha_y.p, (int)H2, g->stream);
      g->launches += 2;
      CK(cudaStreamSynchronize(g->stream));
      CK(cudaGetLastError());
    }
    {  // one-time: the composite rules' table for the raster kernel
      g->comp_lut.reserve(2 * 65536);
      ofdg::launch_composite_luts((uint8_t*)g->comp_lut.p, (uint8_t*)g->comp_lut.p + 65536, g->stream);
      g->launches += 1;
      CK(cudaStreamSynchronize(g->stream));
      CK(cudaGetLastError());
    }
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&g->pipe_uploaded[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->pipe_rendered[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->render_done[i], cudaEventDisableTiming));
    }
    for (cudaEvent_t& e2 : g->chunk_copied) CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming | cudaEventBlockingSync));
    // Host-blob transport: bytes + host widening (176 MB over PCIe, 553 MB through host DRAM per 64-sample step) unless
    // OFDG_TRANSPORT=f32 asks for plain float blobs (403 MB over both). Measured with 8 ranks on one 32-core host
    // (tools/exp_transport_n.sh, profiles/r02_transport_n8.log): bytes 18.6k samples/s for the 8 GPUs together (161 GB/s of host
    // DRAM traffic: the bound), float 15.0k (94 GB/s of inbound PCIe writes: the host's PCIe ingest saturates first) -- bytes
    // win at every rank count.
    if (const char* t = std::getenv("OFDG_TRANSPORT")) g->transport_u8 = std::string(t) != "f32";
    if (const char* t = std::getenv("OFDG_RENDER")) g->split_render = std::string(t) != "fused";
    if (const char* t = std::getenv("OFDG_TEST_PAIR_CAP")) g->pair_cap_limit = std::atoi(t);
    if (const char* t = std::getenv("OFDG_TEST_PHILOX_FG")) g->philox_fg_override = std::atoi(t);
    const char* ov = std::getenv("OFDG_BIN_OVERLAP");
    if (g->split_render && !(ov && std::string(ov) == "0")) {
      // The mask rasterisation of a batch runs on the side stream too, beside the background preparation: it needs the scene
      // and the pair list only, and the two kernels stall on different things (OFDG_RASTER_OVERLAP=0: in line; the per-kernel
      // spans of ofdg_kernel_times then do not overlap). The side stream has the higher priority: the persistent raster
      // blocks must be placed ahead of the preparation's ~20,000 short blocks, or they only start when those have drained.
      // The device-side parameter stream forks it as well (OFDG_PHILOX_RASTER_OVERLAP=0: in line); its look-ahead kernels run on
      // a high-priority stream of their own, or the third concurrent kernel starves them (measured: 109.7k in line, 103.2k
      // forked with the look-ahead at normal priority, 114.3k forked with it at high priority).
      const char* ro = std::getenv("OFDG_RASTER_OVERLAP");
      g->raster_overlap = !(ro && std::string(ro) == "0");
      const char* po = std::getenv("OFDG_PHILOX_RASTER_OVERLAP");
      g->philox_raster_overlap = !(po && std::string(po) == "0");
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&g->bin_stream, cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&g->bin_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->bin_join, cudaEventDisableTiming));
      // cross-batch pipeline (see run_kernels_pipelined): needs the forked raster; OFDG_PIPELINE=0 keeps every batch in line,
      // which is also how per-kernel times are measured (the spans of ofdg_kernel_times then do not overlap)
      const char* pl = std::getenv("OFDG_PIPELINE");
      g->pipeline = g->raster_overlap && !(pl && std::string(pl) == "0");
      g->philox_pipeline = g->pipeline;  // (OFDG_PHILOX_PIPELINE=0: the device-side stream's batches in line)
      if (const char* pp = std::getenv("OFDG_PHILOX_PIPELINE")) g->philox_pipeline = g->pipeline && std::string(pp) != "0";
      if (g->pipeline) {
        const char* pr = std::getenv("OFDG_PREP_PRIORITY");  // "hi": the preparation stream shares the raster's priority
        CK(cudaStreamCreateWithPriority(&g->prep_stream, cudaStreamNonBlocking, (pr && std::string(pr) == "hi") ? hi : 0));
        for (int i = 0; i < 2; ++i) {
          CK(cudaEventCreateWithFlags(&g->set_prep_done[i], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&g->set_raster_done[i], cudaEventDisableTiming));
          CK(cudaEventCreateWithFlags(&g->set_shade_done[i], cudaEventDisableTiming));
        }
      }
    }
    for (ofdg_generator::PrepCtx& c : g->prep_ctx) CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    {
      std::lock_guard<std::mutex> lk(g_live_mu);
      g_live.insert(g.get());
    }
    *out = g.release();
  });
}

void ofdg_destroy(ofdg_generator* g) {
  if (!g) return;
  cudaSetDevice(g->cfg.device);
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live.erase(g);
  }
  cudaDeviceSynchronize();  // render calls on callers' streams may still read the scratch
  g->prep_workers.reset();
  for (ofdg_prepared* p : g->free_scenes) {
    p->scene.release();
    if (p->last_use) cudaEventDestroy(p->last_use);
    delete p;
  }
  g->free_scenes.clear();
  for (ofdg_generator::PrepCtx& c : g->prep_ctx) {
    c.staging.release();
    if (c.stream) cudaStreamDestroy(c.stream);
  }
  g->wf_scratch.release();
  g->reach_host.release();
  if (g->field_stream) cudaStreamDestroy(g->field_stream);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->ph_stream) { cudaStreamSynchronize(g->ph_stream); cudaStreamDestroy(g->ph_stream); }
  for (int i = 0; i < ofdg_generator::kPhSets; ++i) {
    ofdg_generator::PhiloxSet& q = g->ph[i];
    q.scene.release();
    DevBuf* qb[] = {&q.bp, &q.seg_type, &q.seg_x, &q.seg_y, &q.obj_nbp, &q.obj_nseg, &q.ntop, &q.n_deform};
    q.n_deform_host.release();
    for (DevBuf* b : qb) b->release();
    if (q.ready) cudaEventDestroy(q.ready);
    if (q.consumed) cudaEventDestroy(q.consumed);
  }
  DevBuf* bufs[] = {&g->ph_slots, &g->pool, &g->tex_info_dev, &g->fields, &g->fpos_x, &g->falpha_x, &g->fpos_y, &g->falpha_y, &g->mask_raw, &g->mask_warp, &g->field_reach_dev, &g->bg, &g->tile_hits, &g->tile_range, &g->pair_list, &g->pair_masks, &g->pair_ctl, &g->bg_rows, &g->pair_rows, &g->rtab_pos_x, &g->rtab_alpha_x, &g->rtab_pos_y, &g->rtab_alpha_y, &g->out0, &g->out1,
                    &g->outf, &g->ids8, &g->dbg_masks, &g->dbg_id0, &g->dbg_id1, &g->dbg_frames8, &g->dbg_planar};
  for (DevBuf* b : bufs) b->release();
  g->scene.release();
  g->staging.release();
  g->pair_overflow.release();
  for (int i = 0; i < 2; ++i) {
    g->pipe_scene[i].release();
    g->pipe_staging[i].release();
    if (g->pipe_uploaded[i]) cudaEventDestroy(g->pipe_uploaded[i]);
    if (g->pipe_rendered[i]) cudaEventDestroy(g->pipe_rendered[i]);
    if (g->render_done[i]) cudaEventDestroy(g->render_done[i]);
  }
  g->workers.reset();
  for (cudaEvent_t e2 : g->chunk_copied) if (e2) cudaEventDestroy(e2);
  g->out8.release();
  g->host8.release();
  if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
  if (g->bin_stream) { cudaStreamSynchronize(g->bin_stream); cudaStreamDestroy(g->bin_stream); }
  if (g->prep_stream) { cudaStreamSynchronize(g->prep_stream); cudaStreamDestroy(g->prep_stream); }
  for (int i = 0; i < 2; ++i) {
    if (g->set_prep_done[i]) cudaEventDestroy(g->set_prep_done[i]);
    if (g->set_raster_done[i]) cudaEventDestroy(g->set_raster_done[i]);
    if (g->set_shade_done[i]) cudaEventDestroy(g->set_shade_done[i]);
  }
  {
    DevBuf* ab[] = {&g->alt.bg, &g->alt.tile_range, &g->alt.pair_list, &g->alt.pair_masks, &g->alt.pair_ctl, &g->alt.bg_rows, &g->alt.pair_rows};
    for (DevBuf* b : ab) b->release();
  }
  if (g->bin_fork) cudaEventDestroy(g->bin_fork);
  if (g->bin_join) cudaEventDestroy(g->bin_join);
  for (cudaEvent_t e : g->ev_pool) cudaEventDestroy(e);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

// ---- texture pool: textures of any size back to back, each with its foreground view (TexInfo) ---------
namespace {

// Nothing may still be reading the pool (or the look-ahead batch of the device stream) when it changes.
void pool_quiesce(ofdg_generator* g) {
  CK(cudaDeviceSynchronize());
  { g->ph_queue.clear(); g->ph_dirty = true; }
}

// Room for `extra_px` more pixels; keeps what is there.
void pool_grow(ofdg_generator* g, size_t extra_px) {
  const size_t need = (g->pool_px + extra_px) * sizeof(uchar4);
  if (need <= g->pool.cap) return;
  DevBuf bigger;
  bigger.reserve(g->pool_px ? std::max(need, g->pool.cap + g->pool.cap / 2) : need);
  if (g->pool_px) CK(cudaMemcpy(bigger.p, g->pool.p, g->pool_px * sizeof(uchar4), cudaMemcpyDeviceToDevice));
  g->pool.release();
  g->pool = bigger;
}

// Registers the texture whose w x h pixels already sit at pool offset `off`; appends its foreground view when the
// texture is smaller than W x H (Texture::getRandomizedCrop's else branch, DataGenerator.cpp:103-107).
void pool_register(ofdg_generator* g, size_t off, int w, int h) {
  const int W = g->cfg.width, H = g->cfg.height;
  ofdg::TexInfo ti{};
  ti.off = off; ti.w = w; ti.h = h;
  if (w >= W && h >= H) {  // centre crop: crop(w/2-W/2, h/2-H/2, ...) with zoom 1, DataGenerator.cpp:99-102
    ti.fg_base = off + (size_t)(h / 2 - H / 2) * w + (size_t)(w / 2 - W / 2);
    ti.fg_pitch = w;
  } else {
    const size_t P = (size_t)W * H;
    pool_grow(g, P);
    DevBuf tmp, pos, alpha;
    tmp.reserve((size_t)W * h * sizeof(uint32_t));
    pos.reserve((size_t)std::max(W, H) * sizeof(int)); alpha.reserve((size_t)std::max(W, H) * sizeof(double));
    g->launches += ofdg::launch_fg_resize((const uchar4*)g->pool.p + off, w, h, W, H, (uchar4*)g->pool.p + g->pool_px, (uint32_t*)tmp.p,
                                          (int*)pos.p, (double*)alpha.p, g->stream);
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    tmp.release(); pos.release(); alpha.release();
    ti.fg_base = g->pool_px;
    ti.fg_pitch = W;
    g->pool_px += P;
  }
  g->tex_info.push_back(ti);
}

void pool_publish(ofdg_generator* g) {
  g->n_tex = (int)g->tex_info.size();
  g->tex_info_dev.reserve(g->tex_info.size() * sizeof(ofdg::TexInfo));
  CK(cudaMemcpy(g->tex_info_dev.p, g->tex_info.data(), g->tex_info.size() * sizeof(ofdg::TexInfo), cudaMemcpyHostToDevice));
}

void check_texture_size(int n, int w, int h) {
  if (n <= 0) throw ArgError("bad arguments");
  if (w < 2 || h < 2 || w > 32768 || h > 32768) throw ArgError("texture sizes must be in 2..32768");
}

}  // namespace

int ofdg_clear_textures(ofdg_generator* g) {
  return guarded([&] {
    if (!g) throw ArgError("null pointer");
    g->use();
    pool_quiesce(g);
    g->tex_info.clear();
    g->pool_px = 0;
    g->n_tex = 0;
  });
}

int ofdg_add_textures(ofdg_generator* g, const uint8_t* planar, int32_t n, int32_t w, int32_t h) {
  return guarded([&] {
    if (!g || !planar) throw ArgError("bad arguments");
    check_texture_size(n, w, h);
    g->use();
    pool_quiesce(g);
    const size_t plane = (size_t)w * h;
    pool_grow(g, plane * n);
    DevBuf tmp;
    const int chunk = 16;
    tmp.reserve(plane * 3 * chunk);
    const size_t first = g->pool_px;
    for (int i = 0; i < n; i += chunk) {
      const int m = std::min(chunk, n - i);
      CK(cudaMemcpyAsync(tmp.p, planar + (size_t)i * 3 * plane, plane * 3 * m, cudaMemcpyHostToDevice, g->stream));
      ofdg::launch_planar_to_rgbx((const uint8_t*)tmp.p, (uchar4*)g->pool.p + first + (size_t)i * plane, m, w, h, g->stream);
      ++g->launches;
      CK(cudaStreamSynchronize(g->stream));
    }
    tmp.release();
    g->pool_px += plane * n;
    for (int i = 0; i < n; ++i) pool_register(g, first + (size_t)i * plane, w, h);
    pool_publish(g);
  });
}

int ofdg_upload_textures(ofdg_generator* g, const uint8_t* planar, int32_t n, int32_t w, int32_t h) {
  const int rc = ofdg_clear_textures(g);
  return rc ? rc : ofdg_add_textures(g, planar, n, w, h);
}

int ofdg_synth_textures(ofdg_generator* g, int32_t n, int32_t w, int32_t h, uint64_t seed) {
  return guarded([&] {
    if (!g) throw ArgError("bad arguments");
    check_texture_size(n, w, h);
    g->use();
    pool_quiesce(g);
    g->tex_info.clear();
    g->pool_px = 0;
    const size_t plane = (size_t)w * h;
    pool_grow(g, plane * n);
    const int chunk = 64;
    for (int i = 0; i < n; i += chunk) {
      ofdg::launch_synth_textures((uchar4*)g->pool.p + (size_t)i * plane, std::min(chunk, n - i), w, h, seed, i, g->stream);
      ++g->launches;
    }
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    g->pool_px = plane * n;
    for (int i = 0; i < n; ++i) pool_register(g, (size_t)i * plane, w, h);
    pool_publish(g);
  });
}

int ofdg_texture_size(const ofdg_generator* g, int32_t index, int32_t* w, int32_t* h) {
  return guarded([&] {
    if (!g || !w || !h || index < 0 || index >= g->n_tex) throw ArgError("bad arguments");
    *w = g->tex_info[index].w; *h = g->tex_info[index].h;
  });
}

int ofdg_download_texture(ofdg_generator* g, int32_t index, uint8_t* planar_out) {
  return guarded([&] {
    if (!g || !planar_out || index < 0 || index >= g->n_tex) throw ArgError("bad arguments");
    g->use();
    const ofdg::TexInfo& ti = g->tex_info[index];
    const size_t plane = (size_t)ti.w * ti.h;
    g->dbg_planar.reserve(plane * 3);
    ofdg::launch_rgbx_to_planar((const uchar4*)g->pool.p + ti.off, (uint8_t*)g->dbg_planar.p, ti.w, ti.h, g->stream);
    ++g->launches;
    CK(cudaMemcpyAsync(planar_out, g->dbg_planar.p, plane * 3, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
  });
}

// The W x H foreground view of a pool texture as the renderer sees it (parity checks).
int ofdg_download_foreground_view(ofdg_generator* g, int32_t index, uint8_t* planar_out) {
  return guarded([&] {
    if (!g || !planar_out || index < 0 || index >= g->n_tex) throw ArgError("bad arguments");
    g->use();
    const ofdg::TexInfo& ti = g->tex_info[index];
    const int W = g->cfg.width, H = g->cfg.height;
    DevBuf rows;
    rows.reserve((size_t)W * H * sizeof(uchar4));
    CK(cudaMemcpy2D(rows.p, (size_t)W * sizeof(uchar4), (const uchar4*)g->pool.p + ti.fg_base, (size_t)ti.fg_pitch * sizeof(uchar4),
                    (size_t)W * sizeof(uchar4), H, cudaMemcpyDeviceToDevice));
    g->dbg_planar.reserve((size_t)W * H * 3);
    ofdg::launch_rgbx_to_planar((const uchar4*)rows.p, (uint8_t*)g->dbg_planar.p, W, H, g->stream);
    ++g->launches;
    CK(cudaMemcpyAsync(planar_out, g->dbg_planar.p, (size_t)W * H * 3, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    rows.release();
  });
}

int ofdg_set_fields(ofdg_generator* g, const float* fields, int32_t n) {
  return guarded([&] {
    if (!g || !fields || n <= 0) throw ArgError("bad arguments");
    g->use();
    const size_t per = (size_t)2 * 2 * (g->cfg.height + 1) * (g->cfg.width + 1) * sizeof(float);
    g->fields.reserve(per * n);
    CK(cudaMemcpy(g->fields.p, fields, per * n, cudaMemcpyHostToDevice));
    g->n_fields = n;
    // how far the inverse field can move a mask: widens the frame-1 boxes of warped outlines
    const size_t plane2 = (size_t)2 * (g->cfg.height + 1) * (g->cfg.width + 1);
    g->field_reach.assign(n, 0);
    for (int i = 0; i < n; ++i) {
      const float* ifl = fields + ((size_t)i * 2 + 1) * plane2;
      float mx = 0.f;
      for (size_t k = 0; k < plane2; ++k)
        if (std::isfinite(ifl[k])) mx = std::max(mx, std::fabs(ifl[k]));
      g->field_reach[i] = (int)std::ceil(std::min(mx, 1.0e6f));
    }
    g->field_reach_dev.reserve(n * sizeof(int));  // the device-side stream widens the boxes itself
    CK(cudaMemcpy(g->field_reach_dev.p, g->field_reach.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    { g->ph_queue.clear(); g->ph_dirty = true; }
    // CImg linear-resize tables (W+1 -> 2W, H+1 -> 2H) used when the background carries a field
    auto table = [](int len, int nout, std::vector<int>& pos, std::vector<double>& alpha) {
      pos.resize(nout); alpha.resize(nout);
      const double f = nout > 1 ? (len - 1.0) / (nout - 1) : 0;
      double curr = 0, old = 0;
      unsigned q = 0;
      for (int i2 = 0; i2 < nout; ++i2) {
        alpha[i2] = curr - (unsigned int)curr;
        pos[i2] = (int)q;
        old = curr;
        curr = std::min(len - 1.0, curr + f);
        q += (unsigned int)curr - (unsigned int)old;
      }
    };
    std::vector<int> px, py;
    std::vector<double> ax, ay;
    table(g->cfg.width + 1, 2 * g->cfg.width, px, ax);
    table(g->cfg.height + 1, 2 * g->cfg.height, py, ay);
    g->fpos_x.reserve(px.size() * sizeof(int)); g->falpha_x.reserve(ax.size() * sizeof(double));
    g->fpos_y.reserve(py.size() * sizeof(int)); g->falpha_y.reserve(ay.size() * sizeof(double));
    CK(cudaMemcpy(g->fpos_x.p, px.data(), px.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g->falpha_x.p, ax.data(), ax.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g->fpos_y.p, py.data(), py.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g->falpha_y.p, ay.data(), ay.size() * sizeof(double), cudaMemcpyHostToDevice));
  });
}

int ofdg_generate_fields(ofdg_generator* g, uint32_t seed, int32_t n, float* fields_out) {
  if (!g || n <= 0) { g_error = "bad arguments"; return OFDG_ERR_ARG; }
  std::vector<float> host;
  int rc = guarded([&] {
    g->use();
    const size_t per = (size_t)2 * 2 * (g->cfg.height + 1) * (g->cfg.width + 1);
    DevBuf tmp;
    tmp.reserve(per * n * sizeof(float));
    g->launches += ofdg::wf_generate(g->cfg.width, g->cfg.height, seed, n, (float*)tmp.p, g->stream);
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
    host.resize(per * n);
    CK(cudaMemcpy(host.data(), tmp.p, per * n * sizeof(float), cudaMemcpyDeviceToHost));
    tmp.release();
    if (fields_out) std::memcpy(fields_out, host.data(), per * n * sizeof(float));
  });
  if (rc) return rc;
  return ofdg_set_fields(g, host.data(), n);  // install as the generator's pool (reach table, resize tables)
}

int ofdg_reserve_fields(ofdg_generator* g, int32_t total) {
  return guarded([&] {
    if (!g || total <= 0) throw ArgError("bad arguments");
    if (g->n_fields <= 0) throw StateError("ofdg_reserve_fields: install a pool first (ofdg_set_fields / ofdg_generate_fields)");
    if (total <= g->n_fields) return;
    g->use();
    CK(cudaDeviceSynchronize());  // set-up time: nothing may be reading the pool while it moves
    const size_t per = (size_t)2 * 2 * (g->cfg.height + 1) * (g->cfg.width + 1) * sizeof(float);
    DevBuf nf, nr;
    nf.reserve(per * total);
    CK(cudaMemset(nf.p, 0, per * total));  // (new slots read as zero displacement until they are filled)
    CK(cudaMemcpy(nf.p, g->fields.p, per * g->n_fields, cudaMemcpyDeviceToDevice));
    nr.reserve((size_t)total * sizeof(int));
    CK(cudaMemset(nr.p, 0, (size_t)total * sizeof(int)));
    CK(cudaMemcpy(nr.p, g->field_reach_dev.p, (size_t)g->n_fields * sizeof(int), cudaMemcpyDeviceToDevice));
    g->fields.release(); g->field_reach_dev.release();
    g->fields = nf; g->field_reach_dev = nr;
    g->field_reach.resize(total, 0);
    g->n_fields = total;
    { g->ph_queue.clear(); g->ph_dirty = true; }
  });
}

int ofdg_refresh_fields(ofdg_generator* g, uint32_t seed, int32_t first_slot, int32_t n) {
  return guarded([&] {
    if (!g || n <= 0 || first_slot < 0) throw ArgError("bad arguments");
    if (first_slot + n > g->n_fields) throw StateError("ofdg_refresh_fields: slots beyond the pool (size it with ofdg_generate_fields / ofdg_set_fields first)");
    g->use();
    if (!g->field_stream) CK(cudaStreamCreateWithFlags(&g->field_stream, cudaStreamNonBlocking));
    const size_t per = (size_t)2 * 2 * (g->cfg.height + 1) * (g->cfg.width + 1);
    float* dst = (float*)g->fields.p + (size_t)first_slot * per;
    g->launches += ofdg::wf_generate(g->cfg.width, g->cfg.height, seed, n, dst, g->field_stream, &g->wf_scratch);
    int* reach_dev = (int*)g->field_reach_dev.p + first_slot;
    g->launches += ofdg::wf_reach(g->cfg.width, g->cfg.height, dst, n, reach_dev, g->field_stream);
    g->reach_host.reserve((size_t)n * sizeof(int));
    CK(cudaMemcpyAsync(g->reach_host.p, reach_dev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, g->field_stream));
    CK(cudaStreamSynchronize(g->field_stream));
    CK(cudaGetLastError());
    for (int i = 0; i < n; ++i) g->field_reach[first_slot + i] = ((const int*)g->reach_host.p)[i];
  });
}

int ofdg_set_extra_tops(ofdg_generator* g, const ofdg_extra_tops* tops) {
  return guarded([&] {
    if (!g) throw ArgError("null pointer");
    g->use();
    CK(cudaDeviceSynchronize());  // earlier launches may still be writing the previous set
    g->extra = tops ? *tops : ofdg_extra_tops{};
  });
}

int ofdg_render(ofdg_generator* g, const ofdg_task_batch* tasks, float* d_img0, float* d_img1, float* d_flow, void* stream) {
  return guarded([&] {
    if (!g || !tasks || !d_img0 || !d_img1 || !d_flow) throw ArgError("null pointer");
    check_batch(g, tasks->n_tasks);
    g->use();
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    // two scene/staging sets alternate, so that with a caller-provided stream the host flattening of the
    // next call overlaps the kernels of this one; a set is reused only after its last render finished
    const int set = (int)(g->render_calls++ & 1);
    if (g->render_set_used[set]) CK(cudaEventSynchronize(g->render_done[set]));
    flatten_tasks(g, tasks, &g->pipe_flat[set]);
    ensure_scratch(g, tasks->n_tasks);
    upload_scene(g, g->pipe_flat[set], g->pipe_scene[set], g->pipe_staging[set], s);
    run_kernels(g, with_extra_tops(g, make_args(g, g->pipe_scene[set], d_img0, d_img1, d_flow)), s);
    CK(cudaEventRecord(g->render_done[set], s));
    g->render_set_used[set] = true;
    if (!stream) { CK(cudaStreamSynchronize(s)); check_pair_overflow(g); }
  });
}

// Host-blob rendering, pipelined in chunks over five stages that all overlap:
//   draw the parameters (one pool job, sequential: the engines are streams) -> flatten (one pool job per sample)
//   -> [this thread] upload + background preparation + render on the compute stream -> device-to-host copies on
//   the copy stream -> widen the byte frames into the caller's float blobs (pool jobs).
// Unless a sample carries this repository's float augmentation, the frames cross PCIe as bytes (2.3x fewer bytes
// per sample than three float blobs) and host threads widen them -- what the reference does on the host as the
// last step of Process_TaskBucket (/root/reference/src/caffe/DataGenerator.cpp:1228-1244). The flow blob is
// float all the way.
static void render_host_pipelined(ofdg_generator* g, ofdg_params* params, const ofdg_task_batch* tasks, const ofdg_prepared* prepared, int n,
                                  float* h_img0, float* h_img1, float* h_flow) {
  const size_t P = (size_t)g->cfg.width * g->cfg.height;
  bool bytes = g->transport_u8;
  if (params && params->ps->augmentation_enabled()) bytes = false;
  if (prepared && prepared->augmented) bytes = false;
  if (tasks && tasks->augment)
    for (int i = 0; i < n && bytes; ++i) bytes = tasks->augment[i].enabled == 0;
  // 4 samples per chunk: short chunks keep the pipeline's fill and drain short (measured: profiles/README.md)
  int nchunk = std::max(1, std::min(n / 4, 16));
  if (const char* t = std::getenv("OFDG_HOST_CHUNKS")) nchunk = std::max(1, std::min(std::min(std::atoi(t), n), (int)ofdg_generator::kMaxChunks));
  if (bytes) {
    g->out8.reserve((size_t)n * 6 * P);
    g->host8.reserve((size_t)n * 6 * P);
  } else {
    g->out0.reserve((size_t)n * 3 * P * sizeof(float)); g->out1.reserve((size_t)n * 3 * P * sizeof(float));
  }
  g->outf.reserve((size_t)n * 2 * P * sizeof(float));
  if (!g->workers) {
    // widening is bound by host memory bandwidth, not by cores: more threads than this only slow the DMA down
    const int hw = (int)std::thread::hardware_concurrency();
    int threads = std::min(hw / 2, 16);
    if (const char* lw = std::getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU shares the host: an equal share of the cores
      const int local = std::max(1, std::atoi(lw));
      if (local > 1) threads = hw / local - 1;
    }
    if (const char* t = std::getenv("OFDG_HOST_THREADS")) threads = std::atoi(t);
    g->workers.reset(new ofdg::HostPool(std::max(2, std::min(threads, 64)), g->cfg.device));
  }
  ensure_scratch(g, (n + nchunk - 1) / nchunk);
  cudaStream_t A = g->stream, B = g->copy_stream;
  for (int i = 0; i < 2; ++i)  // asynchronous ofdg_render calls share the two scene sets
    if (g->render_set_used[i]) CK(cudaEventSynchronize(g->render_done[i]));

  // ---- host stages: draw + flatten, running ahead of this thread
  static const bool trace = std::getenv("OFDG_TRACE_HOST") != nullptr;  // stage timestamps of every call on stderr
  const auto T0 = std::chrono::steady_clock::now();
  auto now_ms = [T0] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
  double t_ready[ofdg_generator::kMaxChunks] = {}, t_issued[ofdg_generator::kMaxChunks] = {};
  auto& sync = g->host_sync;
  for (int k = 0; k < nchunk; ++k) sync.done[k] = 0;
  sync.error.clear();
  if ((int)g->sample_flat.size() < n) g->sample_flat.resize(n);
  const ofdg::FlattenConfig fc = flatten_config(g);
  auto chunk_of = [n, nchunk](int k) { return (int)((long long)n * k / nchunk); };
  auto fail = [g](const std::string& what) {
    std::lock_guard<std::mutex> lk(g->host_sync.mu);
    if (g->host_sync.error.empty()) g->host_sync.error = what.empty() ? std::string("host stage failed") : what;
    g->host_sync.cv.notify_all();
  };
  auto submit_flatten = [g, fc, fail](ofdg_task_batch one, int sample, int chunk) {  // `one` views exactly one task
    g->workers->submit([g, fc, fail, one, sample, chunk] {
      try {
        g->sample_flat[sample].clear();
        ofdg::flatten(one, fc, g->sample_flat[sample]);
      } catch (const std::exception& e) {
        fail(e.what());
      }
      std::lock_guard<std::mutex> lk(g->host_sync.mu);
      ++g->host_sync.done[chunk];
      g->host_sync.cv.notify_all();
    }, true);
  };
  auto one_task = [](const ofdg_task_batch& all, int i) {
    ofdg_task_batch v = all;
    v.n_tasks = 1;
    v.task_begin = all.task_begin + i;  // blueprint indices stay absolute
    if (v.augment) v.augment += i;
    return v;
  };
  if (prepared) {
    // the scene is already flattened and resident: chunks are windows of its sample array
  } else if (params) {
    ofdg::ParamStream* ps = params->ps.get();
    g->workers->submit([g, ps, nchunk, chunk_of, submit_flatten, one_task, fail] {
      try {
        for (int k = 0; k < nchunk; ++k) {
          ofdg::TaskBatch& tb = g->chunk_tasks[k];
          tb.clear();
          const int t0 = chunk_of(k), t1 = chunk_of(k + 1);
          for (int i = t0; i < t1; ++i) ps->next_task(tb);
          const ofdg_task_batch all = tb.view();
          for (int i = t0; i < t1; ++i) submit_flatten(one_task(all, i - t0), i, k);
        }
      } catch (const std::exception& e) {
        fail(e.what());
      }
    }, true);
  } else {
    for (int k = 0; k < nchunk; ++k)
      for (int i = chunk_of(k); i < chunk_of(k + 1); ++i) submit_flatten(one_task(*tasks, i), i, k);
  }

  size_t uploaded = 0;
  try {
    for (int k = 0; k < nchunk; ++k) {
      const int t0 = chunk_of(k), t1 = chunk_of(k + 1), set = k & 1;
      if (!prepared) {
        {
          std::unique_lock<std::mutex> lk(sync.mu);
          sync.cv.wait(lk, [&] { return sync.done[k] == t1 - t0 || !sync.error.empty(); });
          if (!sync.error.empty()) throw ArgError(sync.error);
        }
        t_ready[k] = now_ms();
        if (k >= 2) CK(cudaEventSynchronize(g->pipe_uploaded[set]));  // the pinned staging area of this set is free again
        upload_scene_parts(g, g->sample_flat.data() + t0, t1 - t0, g->pipe_scene[set], g->pipe_staging[set], A);
        uploaded += g->last_upload_bytes;
        CK(cudaEventRecord(g->pipe_uploaded[set], A));
      }
      float* d0 = bytes ? nullptr : (float*)g->out0.p + (size_t)t0 * 3 * P;
      float* d1 = bytes ? nullptr : (float*)g->out1.p + (size_t)t0 * 3 * P;
      float* df = (float*)g->outf.p + (size_t)t0 * 2 * P;
      ofdg::RenderArgs a = make_args(g, prepared ? prepared->scene : g->pipe_scene[set], d0, d1, df);
      if (prepared) { a.samples += t0; a.batch = t1 - t0; }  // object / shape / vertex indices are absolute within the scene
      if (bytes) a.frames8 = (uint8_t*)g->out8.p + (size_t)t0 * 6 * P;
      run_kernels(g, a, A, !prepared || k == 0);  // a prepared scene's warped masks (mode 9) are made once, by the first chunk
      CK(cudaEventRecord(g->pipe_rendered[set], A));
      CK(cudaStreamWaitEvent(B, g->pipe_rendered[set], 0));
      const size_t c = (size_t)(t1 - t0);
      if (bytes) {
        uint8_t* h8 = (uint8_t*)g->host8.p + (size_t)t0 * 6 * P;
        CK(cudaMemcpyAsync(h8, a.frames8, c * 6 * P, cudaMemcpyDeviceToHost, B));
        CK(cudaEventRecord(g->chunk_copied[k], B));
        CK(cudaMemcpyAsync(h_flow + (size_t)t0 * 2 * P, df, c * 2 * P * sizeof(float), cudaMemcpyDeviceToHost, B));
        std::vector<std::function<void()>> widen;
        widen.reserve(c * 6);
        for (size_t i = 0; i < c; ++i)      // sample t0+i: planes 0-2 are frame 0, planes 3-5 frame 1; one job per plane
          for (int pl = 0; pl < 6; ++pl) {
            float* dst = (pl < 3 ? h_img0 : h_img1) + ((size_t)t0 + i) * 3 * P + (size_t)(pl % 3) * P;
            widen.push_back(g->workers->expand_job(h8 + (i * 6 + pl) * P, dst, P));
          }
        g->workers->submit_after(g->chunk_copied[k], std::move(widen));
      } else {
        CK(cudaMemcpyAsync(h_img0 + (size_t)t0 * 3 * P, d0, c * 3 * P * sizeof(float), cudaMemcpyDeviceToHost, B));
        CK(cudaMemcpyAsync(h_img1 + (size_t)t0 * 3 * P, d1, c * 3 * P * sizeof(float), cudaMemcpyDeviceToHost, B));
        CK(cudaMemcpyAsync(h_flow + (size_t)t0 * 2 * P, df, c * 2 * P * sizeof(float), cudaMemcpyDeviceToHost, B));
      }
      t_issued[k] = now_ms();
    }
  } catch (...) {
    // nothing of this call may still be running (drawing, flattening or writing into the caller's blobs) once it returns
    try { g->workers->wait(); } catch (...) {}
    cudaStreamSynchronize(B);
    cudaStreamSynchronize(A);
    throw;
  }
  CK(cudaStreamSynchronize(B));
  CK(cudaStreamSynchronize(A));
  const double t_copied = now_ms();
  g->workers->wait();
  check_pair_overflow(g);
  if (trace) {
    std::string line = "[ofdg host pipeline] scenes ready / chunk issued (ms):";
    char buf[64];
    for (int k = 0; k < nchunk; ++k) { std::snprintf(buf, sizeof buf, " %.2f/%.2f", t_ready[k], t_issued[k]); line += buf; }
    std::snprintf(buf, sizeof buf, "; copies done %.2f; widened %.2f\n", t_copied, now_ms());
    line += buf;
    std::fputs(line.c_str(), stderr);
  }
  g->last_upload_bytes = uploaded;
  g->last_download_bytes = bytes ? (size_t)n * (6 * P + 2 * P * sizeof(float)) : (size_t)n * 8 * P * sizeof(float);
}

int ofdg_render_host(ofdg_generator* g, const ofdg_task_batch* tasks, float* h_img0, float* h_img1, float* h_flow) {
  return guarded([&] {
    if (!g || !tasks || !h_img0 || !h_img1 || !h_flow) throw ArgError("null pointer");
    check_batch(g, tasks->n_tasks);
    g->use();
    render_host_pipelined(g, nullptr, tasks, nullptr, tasks->n_tasks, h_img0, h_img1, h_flow);
  });
}

int ofdg_generate_host(ofdg_generator* g, ofdg_params* p, int32_t batch, float* h_img0, float* h_img1, float* h_flow) {
  return guarded([&] {
    if (!g || !p || !h_img0 || !h_img1 || !h_flow) throw ArgError("null pointer");
    check_batch(g, batch);
    g->use();
    render_host_pipelined(g, p, nullptr, nullptr, batch, h_img0, h_img1, h_flow);
  });
}

int ofdg_render_debug(ofdg_generator* g, const ofdg_task_batch* tasks, float* h_img0, float* h_img1, float* h_flow,
                      uint8_t* masks, int32_t max_objs, uint32_t* id0, uint32_t* id1, uint8_t* frames8) {
  return guarded([&] {
    if (!g || !tasks) throw ArgError("null pointer");
    check_batch(g, tasks->n_tasks);
    g->use();
    const size_t P = (size_t)g->cfg.width * g->cfg.height, n = tasks->n_tasks;
    g->out0.reserve(n * 3 * P * sizeof(float)); g->out1.reserve(n * 3 * P * sizeof(float)); g->outf.reserve(n * 2 * P * sizeof(float));
    cudaStream_t s = g->stream;
    flatten_tasks(g, tasks);
    ensure_scratch(g, tasks->n_tasks);
    CK(cudaMemsetAsync(g->bg.p, 0, n * 4 * P * sizeof(uchar4), s));
    upload_scene(g, g->flat, g->scene, g->staging, s);
    ofdg::RenderArgs a = make_args(g, g->scene, (float*)g->out0.p, (float*)g->out1.p, (float*)g->outf.p);
    if (masks && max_objs > 0) {
      g->dbg_masks.reserve(n * max_objs * 4 * P);
      CK(cudaMemsetAsync(g->dbg_masks.p, 0, n * max_objs * 4 * P, s));
      a.dbg_masks = (uint8_t*)g->dbg_masks.p;
      a.dbg_max_objs = max_objs;
    }
    if (id0 || id1) {
      g->dbg_id0.reserve(n * P * 4); g->dbg_id1.reserve(n * P * 4);
      a.dbg_id0 = (uint32_t*)g->dbg_id0.p; a.dbg_id1 = (uint32_t*)g->dbg_id1.p;
    }
    if (frames8) {
      g->dbg_frames8.reserve(n * 6 * P);
      a.frames8 = (uint8_t*)g->dbg_frames8.p;
    }
    run_kernels(g, a, s);
    if (h_img0) CK(cudaMemcpyAsync(h_img0, g->out0.p, n * 3 * P * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_img1) CK(cudaMemcpyAsync(h_img1, g->out1.p, n * 3 * P * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_flow) CK(cudaMemcpyAsync(h_flow, g->outf.p, n * 2 * P * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (a.dbg_masks) CK(cudaMemcpyAsync(masks, a.dbg_masks, n * max_objs * 4 * P, cudaMemcpyDeviceToHost, s));
    if (id0) CK(cudaMemcpyAsync(id0, a.dbg_id0, n * P * 4, cudaMemcpyDeviceToHost, s));
    if (id1) CK(cudaMemcpyAsync(id1, a.dbg_id1, n * P * 4, cudaMemcpyDeviceToHost, s));
    if (frames8) CK(cudaMemcpyAsync(frames8, a.frames8, n * 6 * P, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  });
}

int ofdg_debug_background(ofdg_generator* g, const ofdg_task_batch* tasks, uint8_t* planar_out, int32_t* need) {
  return guarded([&] {
    if (!g || !tasks || !planar_out) throw ArgError("null pointer");
    check_batch(g, tasks->n_tasks);
    g->use();
    const size_t P4 = (size_t)g->cfg.width * g->cfg.height * 4, n = tasks->n_tasks;
    cudaStream_t s = g->stream;
    flatten_tasks(g, tasks);
    ensure_scratch(g, tasks->n_tasks);
    CK(cudaMemsetAsync(g->bg.p, 0, n * P4 * sizeof(uchar4), s));
    upload_scene(g, g->flat, g->scene, g->staging, s);
    ofdg::RenderArgs a = make_args(g, g->scene, nullptr, nullptr, nullptr);
    run_kernels(g, a, s);
    g->dbg_planar.reserve(n * 3 * P4);
    ofdg::launch_bg_to_planar((const uchar4*)g->bg.p, (uint8_t*)g->dbg_planar.p, (int)n, 2 * g->cfg.width, 2 * g->cfg.height, s);
    g->launches += n;
    CK(cudaMemcpyAsync(planar_out, g->dbg_planar.p, n * 3 * P4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (need)
      for (size_t i = 0; i < n; ++i) std::memcpy(need + 4 * i, g->flat.samples[i].prep.need, 4 * sizeof(int32_t));
  });
}

int ofdg_debug_composite_luts(ofdg_generator* g, uint8_t* add_lut, uint8_t* sub_lut) {
  return guarded([&] {
    if (!g || !add_lut || !sub_lut) throw ArgError("null pointer");
    g->use();
    g->dbg_planar.reserve(2 * 65536);
    uint8_t* d = (uint8_t*)g->dbg_planar.p;
    ofdg::launch_composite_luts(d, d + 65536, g->stream);
    ++g->launches;
    CK(cudaMemcpyAsync(add_lut, d, 65536, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaMemcpyAsync(sub_lut, d + 65536, 65536, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaStreamSynchronize(g->stream));
    CK(cudaGetLastError());
  });
}

namespace {
void philox_run(ofdg_generator* g, int set, uint64_t seed, uint64_t first_sample, int batch, int augment, int fg_override, cudaStream_t s) {
  using namespace ofdg;
  if (!g->ph_slots.p) {
    SlotSpec specs[kNumSlots];
    fill_mode_table(g->cfg.mode, g->cfg.width, g->cfg.height, specs);
    std::vector<PhiloxSlot> ps(kPhiloxSlots);
    for (int i = 0; i < kNumSlots; ++i) {
      PhiloxSlot q{};
      q.kind = specs[i].kind; q.n_opts = specs[i].n_opts;
      for (int k = 0; k < 4; ++k) q.opts[k] = specs[i].opts[k];
      q.ia = (int)specs[i].a; q.ib = (int)specs[i].b;
      q.a = (float)specs[i].a; q.b = (float)specs[i].b; q.c = (float)specs[i].c; q.d = (float)specs[i].d;
      ps[i] = q;
    }
    g->ph_slots.reserve(ps.size() * sizeof(PhiloxSlot));
    CK(cudaMemcpy(g->ph_slots.p, ps.data(), ps.size() * sizeof(PhiloxSlot), cudaMemcpyHostToDevice));
    double c[100], sn[100];
    for (unsigned st = 0; st < 100; ++st) {  // the unit-circle table of agg::ellipse, from the host's libm (host/flatten.cpp)
      const double angle = double(st) / double(100) * 2.0 * 3.14159265358979323846;
      c[st] = std::cos(angle); sn[st] = std::sin(angle);
    }
    philox_upload_circle(c, sn);
  }
  ofdg_generator::PhiloxSet& q = g->ph[set];
  PhiloxArgs a{};
  a.slots = (const PhiloxSlot*)g->ph_slots.p;
  a.mode = g->cfg.mode; a.W = g->cfg.width; a.H = g->cfg.height;
  a.seed = seed; a.first_sample = first_sample;
  a.batch = batch; a.n_fields = g->cfg.mode == 9 ? g->n_fields : 0; a.fg_override = fg_override; a.augment = augment;
  a.field_reach = (const int*)g->field_reach_dev.p;
  if (!g->pair_overflow.p) {
    g->pair_overflow.reserve(2 * sizeof(int));
    ((volatile int*)g->pair_overflow.p)[0] = 0;
    ((volatile int*)g->pair_overflow.p)[1] = 0;
  }
  a.truncated = (int*)g->pair_overflow.dev + 1;
  a.deform_shape = (int*)q.scene.deform_shape.p; a.deform_field = (int*)q.scene.deform_field.p; a.n_deform = (int*)q.n_deform.p;
  CK(cudaMemsetAsync(q.n_deform.p, 0, sizeof(int), s));
  a.n_tex = g->n_tex; a.tex_info = (const ofdg::TexInfo*)g->tex_info_dev.p;
  a.bp = (ofdg_blueprint*)q.bp.p;
  a.seg_type = (int32_t*)q.seg_type.p; a.seg_x = (float*)q.seg_x.p; a.seg_y = (float*)q.seg_y.p;
  a.obj_nbp = (int*)q.obj_nbp.p; a.obj_nseg = (int*)q.obj_nseg.p; a.n_top = (int*)q.ntop.p;
  a.samples = (FlatSample*)q.scene.samples.p; a.objects = (FlatObject*)q.scene.objects.p;
  a.shapes = (FlatShape*)q.scene.shapes.p; a.verts = (FlatVertex*)q.scene.verts.p;
  g->launches += launch_philox(a, s);
  CK(cudaMemcpyAsync(q.n_deform_host.p, q.n_deform.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  q.scene.batch = batch;
  q.scene.n_deform = 0;  // mode 9: read back by philox_collect once the set is complete
  {
    // The host never sees the boxes of a device-drawn scene. Every object on every tile would be batch * 32 * tiles pairs of
    // 4.4 KB (1.9 GB at batch 64); the modes' scenes have about 300 pairs per 512 x 384 sample, so large batches get 1536 pairs
    // per sample (scaled with the tile count) -- five times the average of a whole batch -- and small ones the full bound.
    // bin_pairs_kernel counts the real pairs: a batch beyond the buffer raises the overflow flag and the call fails loudly.
    const size_t tiles = ofdg::tile_hits_bytes(1, g->cfg.width, g->cfg.height) / ofdg::TILE_HIT_STRIDE;
    const size_t full = (size_t)batch * kPhiloxMaxObj * tiles;
    q.scene.pair_bound = std::min(full, std::max<size_t>((size_t)batch * 1536 * tiles / 192, 16384));
  }
}

// Mode 9: the number of warped outlines of a finished set sizes the deformation pre-pass of its render.
void philox_collect(ofdg_generator* g, int set, cudaStream_t producer) {
  if (g->cfg.mode != 9) return;
  CK(cudaStreamSynchronize(producer));
  g->ph[set].scene.n_deform = *(const int*)g->ph[set].n_deform_host.p;
}

void philox_check(ofdg_generator* g, int batch) {
  using namespace ofdg;
  if (g->n_tex <= 0) throw StateError("no textures uploaded (ofdg_upload_textures / ofdg_synth_textures)");
  if (batch <= 0 || batch > g->cfg.max_batch) throw ArgError("bad batch size");
  if (batch > g->ph_batch) {
    CK(cudaDeviceSynchronize());
    g->ph_queue.clear();
    g->ph_dirty = false;  // (the device was just synchronised)
    const size_t n = batch;
    for (int i = 0; i < ofdg_generator::kPhSets; ++i) {
      ofdg_generator::PhiloxSet& q = g->ph[i];
      q.bp.reserve(n * kPhiloxMaxBp * sizeof(ofdg_blueprint));
      q.seg_type.reserve(n * kPhiloxMaxSeg * sizeof(int32_t)); q.seg_x.reserve(n * kPhiloxMaxSeg * sizeof(float));
      q.seg_y.reserve(n * kPhiloxMaxSeg * sizeof(float));
      q.obj_nbp.reserve(n * kPhiloxMaxObj * sizeof(int)); q.obj_nseg.reserve(n * kPhiloxMaxObj * sizeof(int));
      q.ntop.reserve(n * sizeof(int));
      q.n_deform.reserve(sizeof(int)); q.n_deform_host.reserve(sizeof(int));
      q.scene.deform_shape.reserve(n * kPhiloxMaxObj * kPhiloxMaxShapes * sizeof(int));
      q.scene.deform_field.reserve(n * kPhiloxMaxObj * kPhiloxMaxShapes * sizeof(int));
      q.scene.samples.reserve(n * sizeof(FlatSample));
      q.scene.objects.reserve(n * kPhiloxMaxObj * sizeof(FlatObject));
      q.scene.shapes.reserve(n * kPhiloxMaxObj * kPhiloxMaxShapes * sizeof(FlatShape));
      q.scene.verts.reserve(n * kPhiloxMaxObj * (size_t)kPhiloxMaxVerts * sizeof(FlatVertex));
      q.used = false;
    }
    g->ph_batch = batch;
  }
}
}  // namespace

int ofdg_generate_philox(ofdg_generator* g, uint64_t seed, uint64_t first_sample, int32_t batch, int32_t augment,
                         float* d_img0, float* d_img1, float* d_flow, void* stream) {
  return guarded([&] {
    if (!g || !d_img0 || !d_img1 || !d_flow) throw ArgError("null pointer");
    g->use();
    philox_check(g, batch);
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    int set;
    bool looked_ahead = false;
    std::vector<ofdg_generator::PhNext>& q = g->ph_queue;
    if (!q.empty() && q.front().seed == seed && q.front().first == first_sample && q.front().batch == batch && q.front().augment == augment) {
      set = q.front().set;  // drawn and flattened on the side stream while earlier batches rendered
      q.erase(q.begin());
      CK(cudaStreamWaitEvent(s, g->ph[set].ready, 0));
      philox_collect(g, set, g->ph_stream);
      looked_ahead = true;
    } else {
      if (!q.empty() || g->ph_dirty) CK(cudaStreamSynchronize(g->ph_stream));  // batches nobody asked for may still be written
      q.clear();
      g->ph_dirty = false;
      set = 0;
      if (g->ph[set].used) CK(cudaStreamWaitEvent(s, g->ph[set].consumed, 0));
      philox_run(g, set, seed, first_sample, batch, augment, g->philox_fg_override, s);
      philox_collect(g, set, s);
    }
    ensure_scratch(g, batch);
    // The cross-batch pipeline serves the device-side stream too (OFDG_PHILOX_PIPELINE=0: in line). Measured at batch 64 with the
    // parameter kernel of round 2 (84 us, all lanes working): in line 138.9k samples/s, pipelined 140.2k, pipelined with batches
    // drawn two ahead 143.4k. (With round 1's 169 us kernel the pipelined form was the slower one: 111.1k against 118.6k.)
    const int pset = (g->philox_pipeline && looked_ahead && g->philox_raster_overlap) ? pipeline_set(g, g->ph[set].scene) : -1;
    if (pset >= 0)  // the scene was written on the look-ahead stream: the front end only waits for that, not for the previous batch on s
      run_kernels_pipelined(g, with_extra_tops(g, make_args(g, g->ph[set].scene, d_img0, d_img1, d_flow, pset)), s, pset, g->ph[set].ready);
    else
      run_kernels(g, with_extra_tops(g, make_args(g, g->ph[set].scene, d_img0, d_img1, d_flow)), s, true, g->philox_raster_overlap);
    CK(cudaEventRecord(g->ph[set].consumed, s));
    g->ph[set].used = true;
    // look ahead: the next batches of the same stream, on the side stream, into the sets that are neither being rendered nor
    // waiting in the queue. Two batches ahead (OFDG_PHILOX_DEPTH), a drawn-ahead batch has a whole render step of slack.
    uint64_t next_first = q.empty() ? first_sample + (uint64_t)batch : q.back().first + (uint64_t)batch;
    while ((int)q.size() < g->ph_depth) {
      int nset = -1;
      for (int i = 0; i < ofdg_generator::kPhSets && nset < 0; ++i) {
        bool busy = (i == set);
        for (const ofdg_generator::PhNext& e : q) busy = busy || e.set == i;
        if (!busy) nset = i;
      }
      if (nset < 0) break;
      if (g->ph[nset].used) CK(cudaStreamWaitEvent(g->ph_stream, g->ph[nset].consumed, 0));
      philox_run(g, nset, seed, next_first, batch, augment, g->philox_fg_override, g->ph_stream);
      CK(cudaEventRecord(g->ph[nset].ready, g->ph_stream));
      ofdg_generator::PhNext e;
      e.seed = seed; e.first = next_first; e.batch = batch; e.augment = augment; e.set = nset;
      q.push_back(e);
      next_first += (uint64_t)batch;
    }
    if (!stream) { CK(cudaStreamSynchronize(s)); check_pair_overflow(g); }
  });
}

int ofdg_philox_tasks(ofdg_generator* g, uint64_t seed, uint64_t first_sample, int32_t batch, int32_t augment, ofdg_tasks* out) {
  return guarded([&] {
    if (!g || !out) throw ArgError("null pointer");
    g->use();
    using namespace ofdg;
    philox_check(g, batch);
    CK(cudaDeviceSynchronize());
    { g->ph_queue.clear(); g->ph_dirty = true; }
    philox_run(g, 0, seed, first_sample, batch, augment, 0, g->stream);
    CK(cudaStreamSynchronize(g->stream));
    std::vector<ofdg_blueprint> bp((size_t)batch * kPhiloxMaxBp);
    std::vector<int32_t> st((size_t)batch * kPhiloxMaxSeg);
    std::vector<float> sx(st.size()), sy(st.size());
    std::vector<int> onbp((size_t)batch * kPhiloxMaxObj), onseg((size_t)batch * kPhiloxMaxObj), ntop(batch);
    std::vector<FlatSample> smp(batch);
    CK(cudaMemcpy(bp.data(), g->ph[0].bp.p, bp.size() * sizeof(ofdg_blueprint), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), g->ph[0].seg_type.p, st.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sx.data(), g->ph[0].seg_x.p, sx.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sy.data(), g->ph[0].seg_y.p, sy.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(onbp.data(), g->ph[0].obj_nbp.p, onbp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(onseg.data(), g->ph[0].obj_nseg.p, onseg.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ntop.data(), g->ph[0].ntop.p, batch * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(smp.data(), g->ph[0].scene.samples.p, batch * sizeof(FlatSample), cudaMemcpyDeviceToHost));
    // compact the fixed-stride device arrays into an ordinary task batch
    TaskBatch& tb = out->tb;
    tb.clear();
    for (int s = 0; s < batch; ++s) {
      tb.blueprints.push_back(bp[(size_t)s * kPhiloxMaxBp]);  // background
      for (int k = 0; k < ntop[s]; ++k) {
        const int bp0 = s * kPhiloxMaxBp + 1 + k * kPhiloxMaxShapes, sg0 = s * kPhiloxMaxSeg + k * kPhiloxMaxShapes * 20;
        const int new_bp0 = (int)tb.blueprints.size(), new_sg0 = (int)tb.seg_type.size();
        for (int i = 0; i < onbp[(size_t)s * kPhiloxMaxObj + k]; ++i) {
          ofdg_blueprint b = bp[bp0 + i];
          if (b.seg_count > 0) b.seg_begin = b.seg_begin - sg0 + new_sg0;
          if (b.comp_count > 0) b.comp_begin = b.comp_begin - bp0 + new_bp0;
          if (b.parent >= 0) b.parent = b.parent - bp0 + new_bp0;
          tb.blueprints.push_back(b);
        }
        for (int i = 0; i < onseg[(size_t)s * kPhiloxMaxObj + k]; ++i) { tb.seg_type.push_back(st[sg0 + i]); tb.seg_x.push_back(sx[sg0 + i]); tb.seg_y.push_back(sy[sg0 + i]); }
      }
      tb.task_begin.push_back((int32_t)tb.blueprints.size());
      if (augment) tb.augment.push_back(smp[s].aug);
    }
  });
}

int ofdg_prepare(ofdg_generator* g, const ofdg_task_batch* tasks, ofdg_prepared** out) {
  return guarded([&] {
    if (!g || !tasks || !out) throw ArgError("null pointer");
    check_batch(g, tasks->n_tasks);
    g->use();
    // May run on producer threads beside render calls of another thread: it touches nothing they use (own flatten pool,
    // staging, upload streams), and up to kPrepCtx calls run side by side, each with a context of its own.
    const int n = tasks->n_tasks;
    ofdg_generator::PrepCtx* ctx = nullptr;
    {
      std::unique_lock<std::mutex> lk(g->prepare_mu);
      if (!g->prep_workers) {
        int threads = std::min((int)std::thread::hardware_concurrency() / 2, 8);
        if (const char* lw = std::getenv("LOCAL_WORLD_SIZE")) {  // one process per GPU shares the host
          const int local = std::max(1, std::atoi(lw));
          if (local > 1) threads = std::min(threads, (int)std::thread::hardware_concurrency() / local);
        }
        if (const char* t = std::getenv("OFDG_PREPARE_THREADS")) threads = std::atoi(t);
        g->prep_workers.reset(new ofdg::HostPool(std::max(1, std::min(threads, 32)), g->cfg.device));
      }
      g->prepare_cv.wait(lk, [g] { for (ofdg_generator::PrepCtx& c : g->prep_ctx) if (!c.busy) return true; return false; });
      for (ofdg_generator::PrepCtx& c : g->prep_ctx) if (!c.busy) { ctx = &c; break; }
      ctx->busy = true;
    }
    struct Release {
      ofdg_generator* g; ofdg_generator::PrepCtx* c;
      ~Release() { { std::lock_guard<std::mutex> lk(g->prepare_mu); c->busy = false; } g->prepare_cv.notify_one(); }
    } release{g, ctx};
    if ((int)ctx->flat.size() < n) ctx->flat.resize(n);
    const ofdg::FlattenConfig fc = flatten_config(g);
    // one flatten job per sample (f64 geometry: ellipse 100-gons, curve subdivision, the background's crop parameters)
    ctx->remaining = n;
    ctx->error.clear();
    for (int i = 0; i < n; ++i) {
      ofdg_task_batch one = *tasks;
      one.n_tasks = 1;
      one.task_begin = tasks->task_begin + i;  // blueprint indices stay absolute
      if (one.augment) one.augment += i;
      ofdg::FlatBatch* dst = &ctx->flat[i];
      g->prep_workers->submit([one, fc, dst, ctx] {
        std::string err;
        try {
          dst->clear();
          ofdg::flatten(one, fc, *dst);
        } catch (const std::exception& e) {
          err = e.what();
          if (err.empty()) err = "flatten failed";
        }
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!err.empty() && ctx->error.empty()) ctx->error = err;
        if (--ctx->remaining == 0) ctx->cv.notify_all();
      }, true);
    }
    {
      std::unique_lock<std::mutex> lk(ctx->mu);
      ctx->cv.wait(lk, [ctx] { return ctx->remaining == 0; });
      if (!ctx->error.empty()) throw ArgError(ctx->error);
    }
    std::unique_ptr<ofdg_prepared> p;
    {
      std::lock_guard<std::mutex> lk(g->free_mu);
      if (!g->free_scenes.empty()) {
        p.reset(g->free_scenes.back());
        g->free_scenes.pop_back();
      }
    }
    if (p) {
      if (p->used) CK(cudaEventSynchronize(p->last_use));  // the render that read these buffers last
      p->used = false;
    } else {
      p.reset(new ofdg_prepared);
      CK(cudaEventCreateWithFlags(&p->last_use, cudaEventDisableTiming));
    }
    p->device = g->cfg.device;
    p->owner = g;
    p->augmented = false;
    if (tasks->augment)
      for (int i = 0; i < n; ++i) p->augmented = p->augmented || tasks->augment[i].enabled != 0;
    upload_scene_parts(g, ctx->flat.data(), n, p->scene, ctx->staging, ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));  // the scene is resident (and the staging free) when this returns
    *out = p.release();
  });
}
void ofdg_prepared_destroy(ofdg_prepared* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  {
    std::lock_guard<std::mutex> live(g_live_mu);
    if (p->owner && g_live.count(p->owner)) {
      std::lock_guard<std::mutex> lk(p->owner->free_mu);
      if (p->owner->free_scenes.size() < ofdg_generator::kMaxFreeScenes) {
        p->owner->free_scenes.push_back(p);  // its buffers serve a later ofdg_prepare (which waits for last_use first)
        return;
      }
    }
  }
  if (p->used && p->last_use) cudaEventSynchronize(p->last_use);
  p->scene.release();
  if (p->last_use) cudaEventDestroy(p->last_use);
  delete p;
}
int ofdg_render_prepared(ofdg_generator* g, const ofdg_prepared* p, float* d_img0, float* d_img1, float* d_flow, void* stream) {
  return guarded([&] {
    if (!g || !p || !d_img0 || !d_img1 || !d_flow) throw ArgError("null pointer");
    check_batch(g, p->scene.batch);
    g->use();
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    ensure_scratch(g, p->scene.batch);
    const int set = pipeline_set(g, p->scene);  // the scene has been resident since ofdg_prepare returned
    if (set >= 0) run_kernels_pipelined(g, with_extra_tops(g, make_args(g, p->scene, d_img0, d_img1, d_flow, set)), s, set, nullptr);
    else run_kernels(g, with_extra_tops(g, make_args(g, p->scene, d_img0, d_img1, d_flow)), s);
    if (p->last_use) { CK(cudaEventRecord(p->last_use, s)); p->used = true; }
    if (!stream) { CK(cudaStreamSynchronize(s)); check_pair_overflow(g); }
  });
}

int ofdg_render_prepared_host(ofdg_generator* g, const ofdg_prepared* p, float* h_img0, float* h_img1, float* h_flow) {
  return guarded([&] {
    if (!g || !p || !h_img0 || !h_img1 || !h_flow) throw ArgError("null pointer");
    check_batch(g, p->scene.batch);
    g->use();
    render_host_pipelined(g, nullptr, nullptr, p, p->scene.batch, h_img0, h_img1, h_flow);
  });
}

int ofdg_generate(ofdg_generator* g, ofdg_params* p, int32_t batch, float* d_img0, float* d_img1, float* d_flow, void* stream) {
  if (!g || !p) { g_error = "null pointer"; return OFDG_ERR_ARG; }
  int rc = guarded([&] {
    g->gen_tasks.clear();
    for (int i = 0; i < batch; ++i) p->ps->next_task(g->gen_tasks);
  });
  if (rc) return rc;
  ofdg_task_batch v = g->gen_tasks.view();
  return ofdg_render(g, &v, d_img0, d_img1, d_flow, stream);
}

uint64_t ofdg_launch_count(const ofdg_generator* g) { return g ? g->launches.load() : 0; }
int ofdg_kernel_times(ofdg_generator* g, double* prep_ms, double* render_ms, int32_t* calls) {
  return guarded([&] {
    if (!g) throw ArgError("null pointer");
    g->use();
    CK(cudaDeviceSynchronize());
    check_pair_overflow(g);
    double t[5] = {0, 0, 0, 0, 0};
    for (const ofdg_generator::Span& sp : g->spans) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
      t[sp.kind] += ms;
    }
    g->last_bin_ms = t[3]; g->last_raster_ms = t[4];
    if (prep_ms) *prep_ms = t[0];
    if (render_ms) *render_ms = t[1];
    if (calls) *calls = (int32_t)g->timed_calls;
    g->last_shade_ms = t[2];
    g->spans.clear();
    g->ev_next = 0;
    g->timed_calls = 0;
  });
}
double ofdg_last_shade_ms(const ofdg_generator* g) { return g ? g->last_shade_ms : 0.0; }
int ofdg_last_render_stats(ofdg_generator* g, uint64_t* pairs, uint64_t* prepared_px, uint64_t* source_px) {
  return guarded([&] {
    if (!g) throw ArgError("null pointer");
    g->use();
    CK(cudaDeviceSynchronize());
    int n = 0;
    const void* ctl = g->last_set ? g->alt.pair_ctl.p : g->pair_ctl.p;
    if (ctl) CK(cudaMemcpy(&n, ctl, sizeof(int), cudaMemcpyDeviceToHost));
    if (pairs) *pairs = (uint64_t)n;
    if (prepared_px) *prepared_px = g->last_prep_px;
    if (source_px) *source_px = g->last_prep_src_px;
  });
}
double ofdg_last_bin_ms(const ofdg_generator* g) { return g ? g->last_bin_ms : 0.0; }
double ofdg_last_raster_ms(const ofdg_generator* g) { return g ? g->last_raster_ms : 0.0; }
uint64_t ofdg_last_upload_bytes(const ofdg_generator* g) { return g ? g->last_upload_bytes : 0; }
uint64_t ofdg_last_download_bytes(const ofdg_generator* g) { return g ? g->last_download_bytes : 0; }

}  // extern "C"
