// TEST INFRASTRUCTURE, NOT PRODUCT CODE: C interface over the REFERENCE'S OWN classes.
//
// oracle/ref_build.sh compiles /root/reference/src/caffe/{DataGenerator,WarpFields}.cpp and
// layers/data_generation_layer.cpp untouched, where they lie, against the stand-in headers under
// oracle/shim/ (AGG 2.4, CImg, Caffe are not vendored by the reference and not installed here), and links
// them with this file into oracle/_ref/libofdg_ref.so. Everything the reference's authors wrote -- the 45-engine
// parameter stream, the blueprint realisation, transforms, mask compositing, blits, flow, the worker / queue
// structure, the Caffe layer -- then runs as written; only the third-party arithmetic comes from the shims
// (or from real AGG / CImg when ref_build.sh is pointed at them).
//
// Built with -fno-access-control: the tests re-seed the reference's engines (per-GPU seed offsets) and feed the
// CropGenerator's queue with a fixed pool of warp fields instead of its std::random_device-seeded workers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <random>
#include <string>
#include <vector>

#include "caffe/layers/data_generation_layer.hpp"
#include "caffe/data_generation/DataGenerator.h"

#include "oracle.h"  // oracle_debug (same debug layout as the restatement, so one test harness serves both)

// The reference reads a few blueprint fields it never writes (ObjectBlueprint::is_additive_component,
// do_warpfield_deformation, init_scale, tex_* of foreground objects: SURVEY App. D). Zero-filled allocations make
// those reads deterministic, which is also the convention of ofdg_blueprint ("zero here"). Bound to this library only
// (-Wl,-Bsymbolic-functions); both forms end in malloc/free, so memory may cross into libstdc++ either way.
void* operator new(std::size_t n) {
  void* p = std::calloc(1, n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

namespace {

namespace DG = DataGenerator;
thread_local std::string g_err;

caffe::LayerParameter make_param(int mode, const char* list, int batch, int prefetch, int first, int second, int use_aa) {
  caffe::LayerParameter p;
  p.mutable_data_generation_param()->set_mode(mode);
  p.mutable_data_generation_param()->add_texture_dbases(list ? list : "/dev/null");
  p.mutable_data_generation_param()->set_first_level_threads(first);
  p.mutable_data_generation_param()->set_second_level_threads(second);
  p.mutable_data_generation_param()->set_use_antialiasing(use_aa != 0);
  p.mutable_data_param()->set_batch_size(batch);
  p.mutable_data_param()->set_prefetch(prefetch);
  p.add_top("first-image");
  p.add_top("second-image");
  p.add_top("optical-flow-groundtruth");
  return p;
}

// ---- re-seeding: engine k of the constructor's RNG_SEED++ sequence gets seed_offset + k ----------------------
void reseed(RNG::internal::RNGBase* r, int s) { r->m_mersenne.seed((unsigned)s); }
void reseed(DG::FlyingChairsRandom::Uniform* r, int s) { reseed(&r->m_rng, s); }
template <class T> void reseed(DG::FlyingChairsRandom::Choice<T>* r, int s) { reseed(&r->m_RNG, s); }
void reseed(DG::FlyingChairsRandom::Trigger<DG::FlyingChairsRandom::Uniform>* r, int s) { reseed(&r->m_RNG, s); }
void reseed(DG::FlyingChairsRandom::Gaussian* r, int s) { reseed(&r->m_rng, s); }
void reseed(DG::FlyingChairsRandom::GaussianSq* r, int s) { reseed(&r->m_rng, s); }
void reseed(DG::FlyingChairsRandom::Gaussian3* r, int s) { reseed(&r->m_rng, s); }
void reseed(DG::FlyingChairsRandom::Gaussian4* r, int s) { reseed(&r->m_rng, s); }
void reseed(DG::FlyingChairsRandom::GaussianMeanSigmaRange* r, int s) { reseed(&r->m_rng, s); }

void reseed_all(DG::ObjectParametersGenerator& g, int off) {
  int k = off;  // declaration order of DataGenerator.cpp:1365-1410 (the same in all 13 cases)
  reseed(g.RNG_BgTexID, k++); reseed(g.RNG_BgInitRot, k++); reseed(g.RNG_BgInitTransX, k++); reseed(g.RNG_BgInitTransY, k++);
  reseed(g.RNG_BgRotTrigger, k++); reseed(g.RNG_BgRot, k++); reseed(g.RNG_BgTransX, k++); reseed(g.RNG_BgTransY, k++);
  reseed(g.RNG_BgScaleTrigger, k++); reseed(g.RNG_BgInitScale, k++); reseed(g.RNG_BgScale, k++); reseed(g.RNG_NumberOfFgObjects, k++);
  reseed(g.RNG_ObjType, k++); reseed(g.RNG_ObjTexID, k++); reseed(g.RNG_ObjInitTransX, k++); reseed(g.RNG_ObjInitTransY, k++);
  reseed(g.RNG_ObjTransX, k++); reseed(g.RNG_ObjTransY, k++); reseed(g.RNG_ObjInitRot, k++); reseed(g.RNG_ObjRotTrigger, k++);
  reseed(g.RNG_ObjRot, k++); reseed(g.RNG_ObjInitScale, k++); reseed(g.RNG_ObjScaleTrigger, k++); reseed(g.RNG_ObjScale, k++);
  reseed(g.RNG_ObjTexShiftX, k++); reseed(g.RNG_ObjTexShiftY, k++); reseed(g.RNG_ObjTexRot, k++); reseed(g.RNG_ObjTexZoom, k++);
  reseed(g.RNG_ElliObj_ScaleX, k++); reseed(g.RNG_ElliObj_ScaleY, k++); reseed(g.RNG_PolyObj_spokes, k++); reseed(g.RNG_PolyObj_dphi, k++);
  reseed(g.RNG_PolyObj_r, k++); reseed(g.RNG_PolyObj_ScaleX, k++); reseed(g.RNG_PolyObj_ScaleY, k++); reseed(g.RNG_PolyObj_CurveTrigger, k++);
  reseed(g.RNG_CompObjInitTransX, k++); reseed(g.RNG_CompObjInitTransY, k++); reseed(g.RNG_CompObiNumberOfComponents, k++);
  reseed(g.RNG_ComponentIsAdditive, k++); reseed(g.RNG_ComponentOffset, k++); reseed(g.RNG_ObjIsExtraThin, k++);
  reseed(g.RNG_ObjDeformsNonrigidly, k++); reseed(g.RNG_GenericUniform, k++); reseed(g.RNG_GenericTrigger, k++);
}

// ---- ObjectBlueprint tree <-> flat ofdg_task_batch ----------------------------------------------------------------
struct FlatTasks {
  std::vector<int32_t> task_begin{0};
  std::vector<ofdg_blueprint> bp;
  std::vector<int32_t> seg_type;
  std::vector<float> seg_x, seg_y;
};

size_t push_blueprint(FlatTasks& out, const DG::ObjectBlueprint& b, int parent, bool is_background) {
  ofdg_blueprint f;
  std::memset(&f, 0, sizeof f);
  f.obj_id = b.obj_id;
  // the background's obj_type stays Dummy in the reference (it is never read); the flat record files it as the polygon it is
  f.obj_type = is_background ? (int32_t)OFDG_OBJ_POLYGON : (int32_t)b.obj_type;
  f.init_rot = b.init_rot; f.init_scale = b.init_scale; f.init_trans_x = b.init_trans_x; f.init_trans_y = b.init_trans_y;
  f.rot = b.rot; f.scale = b.scale; f.trans_x = b.trans_x; f.trans_y = b.trans_y;
  f.tex_id = b.tex_id; f.tex_rot = b.tex_rot; f.tex_scale = b.tex_scale; f.tex_shift_x = b.tex_shift_x; f.tex_shift_y = b.tex_shift_y;
  f.ellipse_scale_x = b.ellipse_scale_x; f.ellipse_scale_y = b.ellipse_scale_y;
  f.parent = parent;
  f.is_additive_component = b.is_additive_component;
  f.do_warpfield_deformation = b.do_warpfield_deformation;
  f.field_id = -1;  // not a reference concept: the reference takes whatever crop its queue serves next
  if (!b.polygon_segment_types.empty()) {
    f.seg_begin = (int32_t)out.seg_type.size();
    f.seg_count = (int32_t)b.polygon_segment_types.size();
    for (size_t i = 0; i < b.polygon_segment_types.size(); ++i) {
      out.seg_type.push_back((int32_t)b.polygon_segment_types[i]);
      out.seg_x.push_back(b.polygon_segment_x[i]);
      out.seg_y.push_back(b.polygon_segment_y[i]);
    }
  }
  const size_t idx = out.bp.size();
  out.bp.push_back(f);
  if (!b.composite_component_blueprint_ptrs.empty()) {
    out.bp[idx].comp_begin = (int32_t)out.bp.size();
    out.bp[idx].comp_count = (int32_t)b.composite_component_blueprint_ptrs.size();
    for (const DG::ObjectBlueprint* c : b.composite_component_blueprint_ptrs) push_blueprint(out, *c, (int)idx, false);
  }
  return idx;
}

DG::ObjectBlueprint* make_blueprint(const ofdg_task_batch& t, int idx) {
  const ofdg_blueprint& f = t.blueprints[idx];
  DG::ObjectBlueprint* b = new DG::ObjectBlueprint();
  b->obj_id = f.obj_id;
  b->obj_type = (DG::ObjType_t)f.obj_type;
  b->init_rot = f.init_rot; b->init_scale = f.init_scale; b->init_trans_x = f.init_trans_x; b->init_trans_y = f.init_trans_y;
  b->rot = f.rot; b->scale = f.scale; b->trans_x = f.trans_x; b->trans_y = f.trans_y;
  b->tex_id = f.tex_id; b->tex_rot = f.tex_rot; b->tex_scale = f.tex_scale; b->tex_shift_x = f.tex_shift_x; b->tex_shift_y = f.tex_shift_y;
  b->ellipse_scale_x = f.ellipse_scale_x; b->ellipse_scale_y = f.ellipse_scale_y;
  b->is_additive_component = f.is_additive_component != 0;
  b->do_warpfield_deformation = f.do_warpfield_deformation != 0;
  for (int i = 0; i < f.seg_count; ++i) {
    b->polygon_segment_types.push_back((DG::PolySegmentType_t)t.seg_type[f.seg_begin + i]);
    b->polygon_segment_x.push_back(t.seg_x[f.seg_begin + i]);
    b->polygon_segment_y.push_back(t.seg_y[f.seg_begin + i]);
  }
  for (int i = 0; i < f.comp_count; ++i) b->composite_component_blueprint_ptrs.push_back(make_blueprint(t, f.comp_begin + i));
  return b;
}

struct RefParams {
  std::unique_ptr<DG::ObjectParametersGenerator> gen;
  FlatTasks out;
};

struct RefGenerator {
  std::unique_ptr<DG::DataGenerator> gen;
  DG::RenderCore core;                                   // one worker's RenderCore: never reset between tasks (WorkerThreadLoop)
  std::map<size_t, DG::MovingObjectBase*> objects_map;
  std::unique_ptr<QueueProcessing::QueueProcessor<DG::UnfinishedObjectContainer>> qp;
  std::vector<float> fields;  // n x 2 x 2 x (H+1) x (W+1)
  int n_fields = 0;
  int mode = 1;
};

CImg<float> field_image(const RefGenerator& g, int id, int which) {
  const size_t plane = (size_t)(W + 1) * (H + 1);
  return CImg<float>(g.fields.data() + ((size_t)id * 2 + which) * 2 * plane, W + 1, H + 1, 1, 2);
}

// Queue the crops one task will ask for. policy 0: by the batch's field_id, in the order RealizeObjectBlueprint /
// Process_TaskBucket call get_crop (background, then top-level objects in order; components copy their parent's),
// each crop served once. policy 1: the pool in index order, every crop served three times -- the reference's own
// "reuse_same = 2" consumption (DataGenerator.cpp:1018, WarpFields.cpp:526-531); `cursor` carries on across tasks.
void queue_crops(RefGenerator& g, const ofdg_task_batch& t, int task, int policy, uint64_t* cursor) {
  WarpFields::CropGenerator* cg = g.gen->m_crop_generator_ptr;
  std::queue<std::pair<CImg<float>, CImg<float>>> empty;
  const_cast<int&>(cg->m_reuse_same) = policy == 0 ? 0 : 2;  // (a const member; set at construction in the reference, DataGenerator.cpp:1018)
  if (policy == 0) {
    cg->m_finalized_crops_queue.swap(empty);
    cg->m_reuse_counter = 0;
    for (int i = t.task_begin[task]; i < t.task_begin[task + 1]; ++i) {
      const ofdg_blueprint& f = t.blueprints[i];
      if (f.parent >= 0 || !f.do_warpfield_deformation) continue;
      if (f.field_id < 0 || f.field_id >= g.n_fields) throw std::runtime_error("deformed object without a valid field_id");
      cg->m_finalized_crops_queue.push(std::make_pair(field_image(g, f.field_id, 0), field_image(g, f.field_id, 1)));
    }
  } else {
    int need = 0;
    for (int i = t.task_begin[task]; i < t.task_begin[task + 1]; ++i)
      if (t.blueprints[i].parent < 0 && t.blueprints[i].do_warpfield_deformation) ++need;
    while ((int)cg->m_finalized_crops_queue.size() * 3 < need + 3) {
      const int id = (int)(*cursor % (uint64_t)g.n_fields);
      ++*cursor;
      cg->m_finalized_crops_queue.push(std::make_pair(field_image(g, id, 0), field_image(g, id, 1)));
    }
  }
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

// "shim" or "real" for each third-party dependency this library was built against
const char* ref_describe(void) {
#ifdef OFDG_ORACLE_AGG_SHIM_H_
#define REF_AGG "agg=shim"
#else
#define REF_AGG "agg=real"
#endif
#ifdef OFDG_ORACLE_CIMG_SHIM_H_
#define REF_CIMG "cimg=shim"
#else
#define REF_CIMG "cimg=real"
#endif
  return "reference sources compiled from /root/reference; " REF_AGG " " REF_CIMG " caffe=shim";
}

int ref_width(void) { return W; }
int ref_height(void) { return H; }

// ---- the reference's parameter stream (ObjectParametersGenerator + the commission loop of load_batch) --------------
void* ref_params_create(int mode, int seed_offset) {
  try {
    std::unique_ptr<RefParams> p(new RefParams);
    caffe::LayerParameter lp = make_param(mode, nullptr, 1, 1, 1, 1, 1);
    p->gen.reset(new DG::ObjectParametersGenerator(lp));
    if (seed_offset) reseed_all(*p->gen, seed_offset);
    return p.release();
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_params_destroy(void* h) { delete (RefParams*)h; }
void ref_params_clear(void* h) { ((RefParams*)h)->out = FlatTasks(); }
int ref_params_generate(void* h, int n_tasks) {
  RefParams* p = (RefParams*)h;
  try {
    for (int t = 0; t < n_tasks; ++t) {
      // data_generation_layer.cpp:197-214, statement for statement
      DG::TaskBucket* new_task_ptr = new DG::TaskBucket();
      {
        DG::ObjectBlueprint* b = new DG::ObjectBlueprint();
        b->obj_id = 1;
        p->gen->generateBackground(b);
        new_task_ptr->background_blueprint = b;
      }
      const int fg_objs = p->gen->generateNumberOfFgObjects();
      new_task_ptr->object_blueprints.resize(fg_objs);
      for (int obj_idx = 0; obj_idx < fg_objs; ++obj_idx) {
        DG::ObjectBlueprint* b = new DG::ObjectBlueprint();
        b->obj_id = obj_idx + 10;
        p->gen->generateForegroundObject(b);
        new_task_ptr->object_blueprints[obj_idx] = b;
      }
      push_blueprint(p->out, *new_task_ptr->background_blueprint, -1, true);
      for (DG::ObjectBlueprint* b : new_task_ptr->object_blueprints) push_blueprint(p->out, *b, -1, false);
      p->out.task_begin.push_back((int32_t)p->out.bp.size());
      delete new_task_ptr->background_blueprint;
      for (DG::ObjectBlueprint* b : new_task_ptr->object_blueprints) delete b;
      delete new_task_ptr;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int ref_params_view(void* h, ofdg_task_batch* v) {
  RefParams* p = (RefParams*)h;
  v->n_tasks = (int32_t)p->out.task_begin.size() - 1;
  v->n_blueprints = (int32_t)p->out.bp.size();
  v->n_segments = (int32_t)p->out.seg_type.size();
  v->task_begin = p->out.task_begin.data();
  v->blueprints = p->out.bp.data();
  v->seg_type = p->out.seg_type.data();
  v->seg_x = p->out.seg_x.data();
  v->seg_y = p->out.seg_y.data();
  v->augment = nullptr;
  return 0;
}

// ---- the reference's generator, driven task by task like one of its worker threads ------------------------------------
void* ref_generator_create(int mode, int use_aa, int second_level_threads) {
  try {
    std::unique_ptr<RefGenerator> g(new RefGenerator);
    caffe::LayerParameter lp = make_param(mode, nullptr, 1, 1, 1, second_level_threads, use_aa);
    g->gen.reset(new DG::DataGenerator(lp));
    g->mode = mode;
    g->gen->m_crop_generator_ptr = nullptr;
    if (mode == 9) {  // DataGenerator::Start without the ten std::random_device-seeded producer threads
      g->gen->m_crop_generator_ptr = new WarpFields::CropGenerator(DGEN_WIDTH, DGEN_HEIGHT, 2);
      g->gen->m_crop_generator_ptr->m_running = true;
    }
    // WorkerThreadLoop's per-worker infrastructure (DataGenerator.cpp:1258-1264)
    g->qp.reset(new QueueProcessing::QueueProcessor<DG::UnfinishedObjectContainer>(DG::Process_UnfinishedObjectContainer, false, true,
                                                                                     second_level_threads, true));
    g->qp->SetMaxQueueLength(50).Start();
    return g.release();
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_generator_destroy(void* h) {
  RefGenerator* g = (RefGenerator*)h;
  if (!g) return;
  g->qp.reset();
  if (g->gen->m_crop_generator_ptr) {
    g->gen->m_crop_generator_ptr->m_running = false;
    delete g->gen->m_crop_generator_ptr;
  }
  delete g;
}
// Appends one texture, planar 3 x h x w in the channel order the reference holds after its R<->B swap
// (DataGenerator.cpp:129-131), i.e. exactly what ofdg_upload_textures takes.
int ref_generator_add_texture(void* h, const uint8_t* planar, int w, int hgt) {
  RefGenerator* g = (RefGenerator*)h;
  CImg<unsigned char>* img = new CImg<unsigned char>(planar, w, hgt, 1, 3);
  g->gen->m_random_textures.m_all_textures.push_back(new DG::Texture("memory", img));
  return 0;
}
// TextureCollection's own loader (DataGenerator.cpp:117-149) on a list file; replaces the pool.
int ref_generator_load_list(void* h, const char* path) {
  RefGenerator* g = (RefGenerator*)h;
  try {
    DG::TextureCollection tc(path);
    g->gen->m_random_textures.m_all_textures.swap(tc.m_all_textures);
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int ref_generator_texture_count(void* h) { return (int)((RefGenerator*)h)->gen->m_random_textures.m_all_textures.size(); }
int ref_generator_texture(void* h, int i, uint8_t* planar_out, int* w, int* hgt) {
  RefGenerator* g = (RefGenerator*)h;
  const CImg<unsigned char>& t = *g->gen->m_random_textures.m_all_textures.at(i)->m_texture_ptr;
  *w = t.width(); *hgt = t.height();
  if (planar_out) std::memcpy(planar_out, t.data(), t.size());
  return 0;
}
int ref_generator_set_fields(void* h, const float* fields, int n) {
  RefGenerator* g = (RefGenerator*)h;
  g->n_fields = n;
  g->fields.assign(fields, fields + (size_t)n * 4 * (W + 1) * (H + 1));
  return 0;
}

// Texture::getRandomizedCrop of pool texture tex_id (DataGenerator.cpp:87-109) -> planar 3 x out_h x out_w
int ref_randomized_crop(void* h, int tex_id, int out_w, int out_h, float angle, float zoom, int shift_x, int shift_y, uint8_t* out) {
  RefGenerator* g = (RefGenerator*)h;
  try {
    CImg<unsigned char> r = g->gen->m_random_textures.getTexturePtr(tex_id)->getRandomizedCrop(out_w, out_h, angle, zoom, shift_x, shift_y);
    if (r.width() != out_w || r.height() != out_h || r.spectrum() != 3) throw std::runtime_error("unexpected crop size");
    std::memcpy(out, r.data(), r.size());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// Renders every task of the batch. img0/img1: n x 3 x H x W float, flow: n x 2 x H x W (any may be NULL).
// The outputs come from the reference's own Process_TaskBucket. With `dbg`, the task is realised a second time with the
// same statements as Process_TaskBucket's body (DataGenerator.cpp:1181-1218) so that the objects can be inspected before
// they are deleted; the two passes must agree or the call fails.
int ref_render(void* h, const ofdg_task_batch* tasks, float* img0, float* img1, float* flow, const oracle_debug* dbg, int field_policy) {
  RefGenerator* g = (RefGenerator*)h;
  const size_t P = (size_t)W * H;
  try {
    if (g->gen->m_random_textures.m_all_textures.empty()) throw std::runtime_error("no textures");
    uint64_t cursor = 0;
    for (int t = 0; t < tasks->n_tasks; ++t) {
      DG::TaskBucket task;
      task.task_ID = t;
      const int b0 = tasks->task_begin[t], b1 = tasks->task_begin[t + 1];
      task.background_blueprint = make_blueprint(*tasks, b0);
      for (int i = b0 + 1; i < b1; ++i)
        if (tasks->blueprints[i].parent < 0) task.object_blueprints.push_back(make_blueprint(*tasks, i));

      if (g->mode == 9) queue_crops(*g, *tasks, t, field_policy, &cursor);
      std::queue<std::pair<CImg<float>, CImg<float>>> saved_queue;
      int saved_counter = 0;
      if (g->mode == 9 && dbg) { saved_queue = g->gen->m_crop_generator_ptr->m_finalized_crops_queue; saved_counter = g->gen->m_crop_generator_ptr->m_reuse_counter; }

      g->gen->Process_TaskBucket(&task, g->core, g->objects_map, *g->qp);
      if (task.bad) throw std::runtime_error("reference marked the task bad");
      if (img0) std::memcpy(img0 + (size_t)t * 3 * P, task.result_image0_ptr->data(), 3 * P * sizeof(float));
      if (img1) std::memcpy(img1 + (size_t)t * 3 * P, task.result_image1_ptr->data(), 3 * P * sizeof(float));
      if (flow) std::memcpy(flow + (size_t)t * 2 * P, task.result_flow0_ptr->data(), 2 * P * sizeof(float));

      if (dbg) {
        if (g->mode == 9) { g->gen->m_crop_generator_ptr->m_finalized_crops_queue = saved_queue; g->gen->m_crop_generator_ptr->m_reuse_counter = saved_counter; }
        DG::RenderCore core;
        std::map<size_t, DG::MovingObjectBase*> objects_map;
        DG::ObjectBlueprint* p = task.background_blueprint;
        DG::MovingObjectBackground* bg_obj_ptr = new DG::MovingObjectBackground(p->obj_id);
        bg_obj_ptr->setRawTexture(g->gen->m_random_textures.getTexturePtr(p->tex_id)->getRandomizedCrop(2 * W, 2 * H, p->tex_rot, p->tex_scale,
                                                                                                      p->tex_shift_x, p->tex_shift_y));
        bg_obj_ptr->setMotion(p->rot, p->scale, p->trans_x, p->trans_y);
        if (g->mode == 9 and p->do_warpfield_deformation) {
          CImg<float> warpflow, warpiflow;
          std::tie(warpflow, warpiflow) = g->gen->m_crop_generator_ptr->get_crop();
          warpflow.resize(2 * W, 2 * H, -100, -100, 3);
          warpiflow.resize(2 * W, 2 * H, -100, -100, 3);
          warpflow *= 2.;
          warpiflow *= 2.;
          bg_obj_ptr->setExtraWarpFields(warpflow, warpiflow);
        }
        objects_map[bg_obj_ptr->ID] = bg_obj_ptr;
        g->qp->Give(DG::UnfinishedObjectContainer{bg_obj_ptr});
        for (unsigned int i = 0; i < task.object_blueprints.size(); ++i) {
          DG::MovingObjectBase* base_obj_ptr = g->gen->RealizeObjectBlueprint(task.object_blueprints[i], bg_obj_ptr->m_motion, *g->qp);
          objects_map[base_obj_ptr->ID] = base_obj_ptr;
        }
        g->qp->Finish();
        for (auto it = objects_map.begin(); it != objects_map.end(); ++it)
          if (not core.blitObject(*it->second, g->gen->m_use_AA)) throw std::runtime_error("blitObject failed");
        core.computeFlowImage(objects_map, false);
        core.computeFlowImage(objects_map, true);
        // the two passes must agree
        for (size_t i = 0; i < 3 * P; ++i)
          if ((float)core.frame0.data()[i] != task.result_image0_ptr->data()[i] || (float)core.frame1.data()[i] != task.result_image1_ptr->data()[i])
            throw std::runtime_error("debug pass disagrees with Process_TaskBucket (frames)");
        if (std::memcmp(core.flow0.data(), task.result_flow0_ptr->data(), 2 * P * sizeof(float)) != 0)
          throw std::runtime_error("debug pass disagrees with Process_TaskBucket (flow)");
        if (dbg->id0) for (size_t i = 0; i < P; ++i) dbg->id0[(size_t)t * P + i] = (uint32_t)core.index_image0.data()[i];
        if (dbg->id1) for (size_t i = 0; i < P; ++i) dbg->id1[(size_t)t * P + i] = (uint32_t)core.index_image1.data()[i];
        if (dbg->frames8) {
          std::memcpy(dbg->frames8 + ((size_t)t * 2 + 0) * 3 * P, core.frame0.data(), 3 * P);
          std::memcpy(dbg->frames8 + ((size_t)t * 2 + 1) * 3 * P, core.frame1.data(), 3 * P);
        }
        if (dbg->flow_bw) std::memcpy(dbg->flow_bw + (size_t)t * 2 * P, core.flow1.data(), 2 * P * sizeof(float));
        if (dbg->masks) {
          for (unsigned int k = 0; k < task.object_blueprints.size() && (int)k < dbg->max_objs; ++k) {
            const DG::MovingObjectBase* o = objects_map[task.object_blueprints[k]->obj_id];
            uint8_t* m = dbg->masks + ((size_t)t * dbg->max_objs + k) * 4 * P;
            std::memcpy(m + 0 * P, o->m_masks_AA[0], P);
            std::memcpy(m + 1 * P, o->m_masks_AA[1], P);
            std::memcpy(m + 2 * P, o->m_masks_noAA[0], P);
            std::memcpy(m + 3 * P, o->m_masks_noAA[1], P);
          }
        }
        for (auto it = objects_map.begin(); it != objects_map.end(); ++it) delete it->second;
      }

      delete task.result_image0_ptr;
      delete task.result_image1_ptr;
      delete task.result_flow0_ptr;
      delete task.background_blueprint;
      for (DG::ObjectBlueprint* b : task.object_blueprints) delete b;
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}


// ---- the reference's warp-field producer with an explicit seed ---------------------------------------------------------
// CropGenerator::worker_thread_loop (WarpFields.cpp:540-641) seeds its std::mt19937 from std::random_device and runs on ten
// threads, so its output cannot be reproduced. This is the body of that loop, statement for statement, with the engine
// seeded by the caller and the crops written to `out` (n x 2 x 2 x (H+1) x (W+1)) instead of the queue; the displacer scene,
// DisplacementComposer, FlowField::init_from_DisplacementComposer and clamp_near_zeros are the reference's own code.
int ref_generate_fields(uint32_t seed, int n_fields, float* out) {
  try {
    std::mt19937 mersenne(seed);
    std::uniform_int_distribution<> displacer_type(0, 2);
    std::uniform_real_distribution<> generic_param(-1, 1);
    const int big_size{std::max(W, H) * 3};
    const size_t plane = (size_t)(W + 1) * (H + 1);
    int produced = 0;
    while (produced < n_fields) {
      WarpFields::DisplacementComposer dc(big_size, big_size);
      const int spacing{200};
      const int isosceles_spacing{(int)(spacing / 2. * std::sqrt(3.))};
      const int rows{(dc.get_H() + isosceles_spacing - 1) / isosceles_spacing};
      const int cols{(dc.get_W()) / spacing};
      for (int yidx = 0; yidx < rows; ++yidx) {
        for (int xidx = 0; xidx < cols; ++xidx) {
          const int x = xidx * spacing + (yidx % 2 == 1 ? spacing / 2 : 0) + spacing / 2;
          const int y = yidx * isosceles_spacing + spacing / 2;
          // The reference draws these inside constructor argument lists (WarpFields.cpp:579-601), whose evaluation order C++
          // leaves unspecified; the draws are i.i.d., so any order is "the reference". Fixed here to left-to-right so that the
          // restatement (oracle/warpfields.cpp) can be compared value for value.
          auto g = [&]() { return generic_param(mersenne); };
          WarpFields::Displacers::DisplacerBase* displacer_ptr{nullptr};
          switch (displacer_type(mersenne)) {
            case 0: { const double a = g() * 3e-4, b = g() * 3e-4; displacer_ptr = new WarpFields::Displacers::Translation(a, b); break; }
            case 1: { const double a = x + g() * 10, b = y + g() * 10, c = g() * M_PI * 2e-6; displacer_ptr = new WarpFields::Displacers::Rotation(a, b, c); break; }
            case 2: { const double a = x + g() * 10, b = y + g() * 10, c = 1 + g() * 2e-6; displacer_ptr = new WarpFields::Displacers::Zoom(a, b, c); break; }
          }
          const double s0 = x + g() * 10, s1 = y + g() * 10, s2 = 50 + g() * 20, s3 = 50 + g() * 20, s4 = g() * M_PI;
          WarpFields::Supports::SupportBase* support_ptr = new WarpFields::Supports::Gaussian2D(s0, s1, s2, s3, s4);
          dc.add_displacer(displacer_ptr).with_support(support_ptr);
        }
      }
      WarpFields::FlowField ff;
      ff.init_from_DisplacementComposer(dc).clamp_near_zeros();
      const CImg<float> flow = ff.get_flow();
      const CImg<float> iflow = ff.get_iflow();
      for (int y = H / 4; y < big_size - 5 * H / 4 && produced < n_fields; y += H / 3) {
        for (int x = W / 4; x < big_size - 5 * W / 4 && produced < n_fields; x += W / 3) {
          CImg<float> crop = flow.get_crop(x, y, x + W, y + H);
          CImg<float> icrop = iflow.get_crop(x, y, x + W, y + H);
          float* dst = out + (size_t)produced * 4 * plane;
          std::memcpy(dst, crop.data(), 2 * plane * sizeof(float));
          std::memcpy(dst + 2 * plane, icrop.data(), 2 * plane * sizeof(float));
          ++produced;
        }
      }
    }
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// ---- the reference's Caffe layer, end to end (bench.py --impl reference / cpu_baseline) ------------------------------
struct RefLayer {
  std::shared_ptr<caffe::Layer<float>> layer;
  caffe::Blob<float> tops[3];
  std::vector<caffe::Blob<float>*> bottom, top;
};
void* ref_layer_create(int mode, const char* texture_list, int batch, int prefetch, int first_level_threads, int second_level_threads, int use_aa) {
  try {
    std::unique_ptr<RefLayer> l(new RefLayer);
    caffe::LayerParameter lp = make_param(mode, texture_list, batch, prefetch, first_level_threads, second_level_threads, use_aa);
    l->layer = caffe::LayerRegistry<float>::CreateLayer(lp);  // REGISTER_LAYER_CLASS(DataGeneration), data_generation_layer.cpp:299
    for (int i = 0; i < 3; ++i) l->top.push_back(&l->tops[i]);
    l->layer->SetUp(l->bottom, l->top);
    return l.release();
  } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
int ref_layer_forward(void* h, float* img0, float* img1, float* flow) {
  RefLayer* l = (RefLayer*)h;
  try {
    l->layer->Forward(l->bottom, l->top);
    if (img0) std::memcpy(img0, l->tops[0].cpu_data(), sizeof(float) * l->tops[0].count());
    if (img1) std::memcpy(img1, l->tops[1].cpu_data(), sizeof(float) * l->tops[1].count());
    if (flow) std::memcpy(flow, l->tops[2].cpu_data(), sizeof(float) * l->tops[2].count());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
int ref_layer_top_shape(void* h, int i, int* shape4) {
  RefLayer* l = (RefLayer*)h;
  for (int k = 0; k < 4; ++k) shape4[k] = l->tops[i].shape().at(k);
  return 0;
}
void ref_layer_destroy(void* h) { delete (RefLayer*)h; }

}  // extern "C"
