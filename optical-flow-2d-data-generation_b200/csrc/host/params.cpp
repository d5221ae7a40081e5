// See params.hpp. Citations: /root/reference/src/caffe/DataGenerator.cpp (DG.cpp),
// /root/reference/include/caffe/data_generation/SimpleRandom.h (SR.h).
#include "params.hpp"

#include <algorithm>

#include <cmath>
#include <limits>
#include <stdexcept>

namespace ofdg {

void fill_mode_table(int mode, int W, int H, SlotSpec out[kNumSlots]) {
  const double pi = 3.14159265358979323846;  // agg::pi
  const int int_max = std::numeric_limits<int>::max();
  (void)pi; (void)int_max; (void)W; (void)H;
  auto set = [&](int slot, int kind, double a = 0, double b = 0, double c = 0, double d = 0) {
    SlotSpec s;
    s.kind = kind;
    if (kind == CHOICE_INT || kind == CHOICE_TYPE) {
      s.n_opts = (int)a;
      s.opts[0] = (int)b; s.opts[1] = (int)c; s.opts[2] = (int)d;
    } else {
      s.a = a; s.b = b; s.c = c; s.d = d;
    }
    out[slot] = s;
  };
#define S(slot, kind, ...) set(slot, kind, __VA_ARGS__);
  switch (mode) {
#include "mode_tables.inc"
    default:
      throw std::runtime_error("BAD MODE");
  }
#undef S
}

const char* slot_name(int slot) {
  static const char* names[kNumSlots] = {
#define OFDG_SLOT_NAMES
#define N(i, n) #n,
#include "mode_tables.inc"
#undef N
#undef OFDG_SLOT_NAMES
  };
  return (slot >= 0 && slot < kNumSlots) ? names[slot] : "?";
}

// ---------------------------------------------------------------------------------------------
Engine& Engine::operator=(Engine&& o) noexcept {
  draws = o.draws; spec_ = o.spec_; mt_ = o.mt_; int_ = o.int_; real_ = o.real_; normal_ = o.normal_;
  fa_ = o.fa_; fb_ = o.fb_; fc_ = o.fc_; fd_ = o.fd_;
  cur_ = std::move(o.cur_); pos_ = o.pos_; ready_ = std::move(o.ready_); mu_ = std::move(o.mu_);
  produced_.store(o.produced_.load()); consumed_ = o.consumed_; shared_ = o.shared_;
  return *this;
}

Engine::Engine(const SlotSpec& spec, int seed) : spec_(spec), mt_((uint32_t)seed), mu_(new std::mutex) {
  // The reference's constructors take float a/b (DG.h:301-363); narrow here, widen where
  // the std:: distribution wants double (SR.h:98-102).
  fa_ = (float)spec.a; fb_ = (float)spec.b; fc_ = (float)spec.c; fd_ = (float)spec.d;
  switch (spec.kind) {
    case UINT:
      int_ = std::uniform_int_distribution<int>((int)spec.a, (int)spec.b);  // SR.h:78-82
      break;
    case CHOICE_INT:
    case CHOICE_TYPE:
      int_ = std::uniform_int_distribution<int>(0, spec.n_opts - 1);  // DG.cpp:853-856
      break;
    case UREAL:
    case TRIGGER:
      real_ = std::uniform_real_distribution<double>(fa_, fb_);  // SR.h:98-102
      break;
    default:
      normal_ = std::normal_distribution<float>(0.f, 1.f);  // SR.h:133-135 with (0, 1, seed)
      break;
  }
}

// DG.cpp:828-831
static inline float base_gauss(float a, float b, float input, float normalize) {
  float sample{input * ((b + a) / 2.f - a) / normalize + (b + a) / 2.f};
  return ((a <= sample && sample <= b) ? sample : (b + a) / 2.);
}

Engine::Value Engine::produce() {
  Value v;
  switch (spec_.kind) {
    case UREAL:
      v.f = (float)real_(mt_);  // SR.h:104-105 (double narrowed by the float return type)
      return v;
    case GAUSS_SQ: {             // DG.cpp:886-890
      float tmp = normal01();
      tmp = ((tmp > 0) ? std::pow(tmp, 2) : -std::pow(tmp, 2));
      v.f = base_gauss(fa_, fb_, tmp, 6);
      return v;
    }
    case GAUSS_3: {              // DG.cpp:897-900
      float tmp = std::pow(normal01(), 3);
      v.f = base_gauss(fa_, fb_, tmp, 10);
      return v;
    }
    case GAUSS_4: {              // DG.cpp:907-911
      float tmp = normal01();
      tmp = ((tmp > 0) ? std::pow(tmp, 4) : -std::pow(tmp, 4));
      v.f = base_gauss(fa_, fb_, tmp, 15);
      return v;
    }
    case GAUSS_MSR: {            // DG.cpp:918-921: a, b, mean = c, sigma = d
      float tmp = normal01() * fd_ + fc_;
      v.f = (((fa_ <= tmp) && (tmp <= fb_)) ? tmp : fc_);
      return v;
    }
    case UINT:
      v.i = int_(mt_);
      return v;
    case CHOICE_INT:
    case CHOICE_TYPE:
      v.i = spec_.opts[int_(mt_)];  // DG.cpp:859-861
      return v;
    default: {                   // TRIGGER, DG.cpp:846-849
      float u = (float)real_(mt_);
      v.i = u < fc_ ? 1 : 0;
      return v;
    }
  }
}

void Engine::fill(size_t count) {
  std::lock_guard<std::mutex> l(*mu_);
  std::vector<Value> chunk;
  chunk.reserve(count);
  for (size_t i = 0; i < count; ++i) chunk.push_back(produce());
  ready_.push_back(std::move(chunk));
  produced_.fetch_add(count, std::memory_order_relaxed);
}

Engine::Value Engine::next_slow() {
  std::lock_guard<std::mutex> l(*mu_);  // (waits for a fill in progress: the engine's state is its alone)
  while (!ready_.empty()) {
    cur_ = std::move(ready_.front());
    ready_.pop_front();
    pos_ = 0;
    if (!cur_.empty()) { ++consumed_; return cur_[pos_++]; }
  }
  cur_.clear();
  pos_ = 0;
  return produce();
}

float Engine::real() {
  ++draws;
  if (spec_.kind != UREAL && spec_.kind < GAUSS_SQ) throw std::logic_error("Engine::real on a non-real slot");
  return next().f;
}

int Engine::integer() {
  ++draws;
  if (spec_.kind != UINT && spec_.kind != CHOICE_INT && spec_.kind != CHOICE_TYPE) throw std::logic_error("Engine::integer on a non-integer slot");
  return next().i;
}

bool Engine::trigger() {  // DG.cpp:846-849
  ++draws;
  if (spec_.kind != TRIGGER) throw std::logic_error("Engine::trigger on a non-trigger slot");
  return next().i != 0;
}

// ---------------------------------------------------------------------------------------------
MiniPool::MiniPool(int threads) {
  for (int i = 0; i < threads; ++i) workers_.emplace_back([this] { loop(); });
}
MiniPool::~MiniPool() {
  {
    std::lock_guard<std::mutex> l(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (std::thread& t : workers_) t.join();
}
void MiniPool::loop() {
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> l(mu_);
      cv_work_.wait(l, [this] { return stop_ || !queue_.empty(); });
      if (queue_.empty()) return;
      job = std::move(queue_.front());
      queue_.pop_front();
    }
    job();
    std::lock_guard<std::mutex> l(mu_);
    if (--pending_ == 0) cv_done_.notify_all();
  }
}
void MiniPool::start(std::vector<std::function<void()> > jobs) {
  if (jobs.empty()) return;
  {
    std::lock_guard<std::mutex> l(mu_);
    pending_ += jobs.size();
    for (std::function<void()>& j : jobs) queue_.push_back(std::move(j));
  }
  cv_work_.notify_all();
}
void MiniPool::wait() {
  std::unique_lock<std::mutex> l(mu_);
  cv_done_.wait(l, [this] { return pending_ == 0; });
}

// ---------------------------------------------------------------------------------------------
void TaskBatch::clear() {
  task_begin.assign(1, 0);
  blueprints.clear();
  seg_type.clear();
  seg_x.clear();
  seg_y.clear();
  augment.clear();
}

ofdg_task_batch TaskBatch::view() const {
  ofdg_task_batch v;
  v.n_tasks = n_tasks();
  v.n_blueprints = (int32_t)blueprints.size();
  v.n_segments = (int32_t)seg_type.size();
  v.task_begin = task_begin.data();
  v.blueprints = blueprints.data();
  v.seg_type = seg_type.data();
  v.seg_x = seg_x.data();
  v.seg_y = seg_y.data();
  v.augment = (augment.size() == (size_t)n_tasks() && !augment.empty()) ? augment.data() : nullptr;
  return v;
}

static ofdg_blueprint blank_blueprint() {
  ofdg_blueprint b{};  // zero everything the reference leaves uninitialised
  b.obj_type = OFDG_OBJ_DUMMY;
  b.parent = -1;
  b.field_id = -1;
  return b;
}

ParamStream::ParamStream(int mode, int W, int H, int seed_offset, int n_fields, int fg_override)
    : mode_(mode), W_(W), H_(H), n_fields_(n_fields), fg_override_(fg_override) {
  SlotSpec specs[kNumSlots];
  fill_mode_table(mode, W, H, specs);
  for (int i = 0; i < kNumSlots; ++i) eng_[i] = Engine(specs[i], seed_offset + i);  // RNG_SEED++, DG.cpp:1360 (move-assigned)
  // augmentation engines (this repository's own spec): seeded far away from every rank's 45 reference seeds (ranks are strided by
  // 45: seeds seed_offset + 45.. would be the next rank's first engines, and the streams of neighbouring GPUs would be correlated)
  for (int i = 0; i < 5; ++i) aug_eng_[i] = std::mt19937(0x40000000u + (uint32_t)seed_offset + (uint32_t)i);
}

// CropGenerator::get_crop serves every crop reuse_same+1 = 3 times before popping it
// (WarpFields.cpp:516-538, DG.cpp:1018). The pool is injected, so ids cycle through it.
int ParamStream::next_field() {
  if (n_fields_ <= 0) return -1;
  int id = (int)((field_draws_ / 3) % (uint64_t)n_fields_);
  ++field_draws_;
  return id;
}

void ParamStream::background(ofdg_blueprint& b) {  // DG.cpp:2105-2143
  b.rot = (eng_[BgRotTrigger].trigger() ? eng_[BgRot].real() : 0.);
  b.scale = (eng_[BgScaleTrigger].trigger() ? eng_[BgScale].real() : 1.);
  float pre_transx = eng_[BgTransX].real();
  float pre_transy = eng_[BgTransY].real();
  b.trans_x = std::cos(-b.rot) * pre_transx - std::sin(-b.rot) * pre_transy;
  b.trans_y = std::sin(-b.rot) * pre_transx + std::cos(-b.rot) * pre_transy;
  b.tex_id = eng_[BgTexID].integer();
  b.tex_rot = eng_[BgInitRot].real();
  b.tex_scale = eng_[BgInitScale].real();
  b.tex_shift_x = eng_[BgInitTransX].integer();
  b.tex_shift_y = eng_[BgInitTransY].integer();
  b.do_warpfield_deformation = eng_[ObjDeformsNonrigidly].trigger();
  if (mode_ == 9 && b.do_warpfield_deformation) b.field_id = next_field();  // DG.cpp:1194-1196
}

void ParamStream::common_prefix(ofdg_blueprint& b, bool is_component, bool allow_redraw) {
  // Type: a component is pre-marked Composite and redraws until it is not (DG.cpp:2326-2332, 2441-2444).
  do {
    b.obj_type = eng_[ObjType].integer();
  } while (allow_redraw && is_component && b.obj_type == OFDG_OBJ_COMPOSITE);
  b.init_rot = eng_[ObjInitRot].real();
  b.init_trans_x = eng_[ObjInitTransX].real();
  b.init_trans_y = eng_[ObjInitTransY].real();
  b.rot = (eng_[ObjRotTrigger].trigger() ? eng_[ObjRot].real() : 0.);
  b.scale = (eng_[ObjScaleTrigger].trigger() ? eng_[ObjScale].real() : 1.);
  b.trans_x = eng_[ObjTransX].real();
  b.trans_y = eng_[ObjTransY].real();
  b.tex_id = eng_[ObjTexID].integer();
  if (mode_ == 9) b.do_warpfield_deformation = eng_[ObjDeformsNonrigidly].trigger();  // DG.cpp:2619
}

void ParamStream::ellipse_params(ofdg_blueprint& b) {  // e.g. DG.cpp:2460-2461
  b.ellipse_scale_x = eng_[ElliObj_ScaleX].real() * 50;
  b.ellipse_scale_y = eng_[ElliObj_ScaleY].real() * 50;
}

void ParamStream::polygon_params(TaskBatch& out, ofdg_blueprint& b, bool curves) {
  const double pi = 3.14159265358979323846;
  b.seg_begin = (int32_t)out.seg_type.size();
  if (mode_ == 1) {  // axis-aligned box, DG.cpp:2163-2183
    const float radius = eng_[PolyObj_r].real();
    const float xscale = radius * eng_[PolyObj_ScaleX].real();
    const float yscale = radius * eng_[PolyObj_ScaleY].real();
    const float xs[4] = {xscale, xscale, -xscale, -xscale};
    const float ys[4] = {-yscale, yscale, yscale, -yscale};
    for (int i = 0; i < 4; ++i) {
      out.seg_x.push_back(xs[i]);
      out.seg_y.push_back(ys[i]);
      out.seg_type.push_back(i == 0 ? OFDG_SEG_DUMMY : OFDG_SEG_LINE);
    }
    b.seg_count = 4;
    return;
  }
  // star polygon, DG.cpp:2469-2495
  const unsigned int spokes = static_cast<unsigned int>(eng_[PolyObj_spokes].integer());
  std::vector<float> phi(spokes), r(spokes);
  for (unsigned int i = 0; i < spokes; ++i) {
    phi[i] = (i * 360. / spokes + eng_[PolyObj_dphi].real()) * pi / 180.;
    r[i] = eng_[PolyObj_r].real();
  }
  const float xscale = eng_[PolyObj_ScaleX].real();
  const float yscale = eng_[PolyObj_ScaleY].real();
  for (unsigned int i = 0; i < spokes; ++i) {
    out.seg_x.push_back(xscale * r[i] * std::cos(phi[i]));
    out.seg_y.push_back(yscale * r[i] * std::sin(phi[i]));
    out.seg_type.push_back(OFDG_SEG_LINE);
  }
  int32_t* types = out.seg_type.data() + b.seg_begin;
  types[0] = OFDG_SEG_DUMMY;
  for (unsigned int i = 1; i < spokes; ++i) {
    if (curves && (i < spokes - 1) && eng_[PolyObj_CurveTrigger].trigger()) {
      types[i] = OFDG_SEG_CURVE3;
      types[i + 1] = OFDG_SEG_DUMMY;
      ++i;
    } else {
      types[i] = OFDG_SEG_LINE;
    }
  }
  b.seg_count = (int32_t)spokes;
}

static void copy_placement(ofdg_blueprint& c, const ofdg_blueprint& b) {  // DG.cpp:2556-2563
  c.init_rot = b.init_rot;
  c.init_trans_x = b.init_trans_x;
  c.init_trans_y = b.init_trans_y;
  c.rot = b.rot;
  c.scale = b.scale;
  c.trans_x = b.trans_x;
  c.trans_y = b.trans_y;
}

static void shrink(TaskBatch& out, ofdg_blueprint& c, double f) {
  if (c.obj_type == OFDG_OBJ_ELLIPSE) {
    c.ellipse_scale_x *= f;
    c.ellipse_scale_y *= f;
  } else if (c.obj_type == OFDG_OBJ_POLYGON) {
    for (int si = 0; si < c.seg_count; ++si) {
      out.seg_x[c.seg_begin + si] *= f;
      out.seg_y[c.seg_begin + si] *= f;
    }
  } else {
    throw std::runtime_error("Bad component object type");
  }
}

void ParamStream::composite_parts(TaskBatch& out, size_t idx) {  // DG.cpp:2384-2426, 2549-2591
  const unsigned int parts = eng_[CompObiNumberOfComponents].integer();
  out.blueprints[idx].comp_begin = (int32_t)out.blueprints.size();
  out.blueprints[idx].comp_count = (int32_t)parts;
  for (unsigned int part_idx = 0; part_idx < parts; ++part_idx) {
    size_t ci = out.blueprints.size();
    out.blueprints.push_back(blank_blueprint());
    out.blueprints[ci].obj_type = OFDG_OBJ_COMPOSITE;
    foreground(out, ci, true);
    ofdg_blueprint& c = out.blueprints[ci];
    const ofdg_blueprint& b = out.blueprints[idx];
    c.parent = (int32_t)idx;
    copy_placement(c, b);
    if (part_idx == 0) {
      c.is_additive_component = true;
    } else {
      c.init_rot = eng_[ObjInitRot].real();
      c.init_trans_x += eng_[ComponentOffset].real();
      c.init_trans_y += eng_[ComponentOffset].real();
      shrink(out, c, 0.2);
      c.is_additive_component = eng_[ComponentIsAdditive].trigger();
    }
    if (mode_ == 9) {  // DG.cpp:2756
      c.do_warpfield_deformation = b.do_warpfield_deformation;
      c.field_id = b.field_id;
    }
  }
}

void ParamStream::outline_parts(TaskBatch& out, size_t idx) {  // DG.cpp:2504-2547, 2668-2713
  out.blueprints[idx].comp_begin = (int32_t)out.blueprints.size();
  out.blueprints[idx].comp_count = 2;
  size_t i1 = out.blueprints.size();
  out.blueprints.push_back(blank_blueprint());
  out.blueprints[i1].obj_type = OFDG_OBJ_COMPOSITE;
  foreground(out, i1, true);
  {
    ofdg_blueprint& c1 = out.blueprints[i1];
    const ofdg_blueprint& b = out.blueprints[idx];
    c1.parent = (int32_t)idx;
    copy_placement(c1, b);
    c1.is_additive_component = true;
    if (mode_ == 9) {
      c1.do_warpfield_deformation = b.do_warpfield_deformation;
      c1.field_id = b.field_id;
    }
  }
  // c2 = copy of c1, with its own copy of the polygon segments
  size_t i2 = out.blueprints.size();
  out.blueprints.push_back(out.blueprints[i1]);
  ofdg_blueprint& c2 = out.blueprints[i2];
  const ofdg_blueprint& c1 = out.blueprints[i1];
  const ofdg_blueprint& b = out.blueprints[idx];
  if (c1.obj_type == OFDG_OBJ_POLYGON) {
    c2.seg_begin = (int32_t)out.seg_type.size();
    for (int si = 0; si < c1.seg_count; ++si) {
      out.seg_type.push_back(out.seg_type[c1.seg_begin + si]);
      out.seg_x.push_back(out.seg_x[c1.seg_begin + si]);
      out.seg_y.push_back(out.seg_y[c1.seg_begin + si]);
    }
  }
  if (c1.obj_type == OFDG_OBJ_ELLIPSE) {
    if (eng_[GenericTrigger].trigger()) {
      c2.init_trans_x = b.init_trans_x + eng_[CompObjInitTransX].real();
      c2.init_trans_y = b.init_trans_y + eng_[CompObjInitTransY].real();
    } else {
      c2.init_trans_x = b.init_trans_x;
      c2.init_trans_y = b.init_trans_y;
      c2.ellipse_scale_x *= 0.9;
      c2.ellipse_scale_y *= 0.9;
    }
  } else {
    c2.init_trans_x = b.init_trans_x;
    c2.init_trans_y = b.init_trans_y;
    shrink(out, c2, 0.9);
  }
  c2.scale = b.scale;
  c2.rot = b.rot;
  c2.trans_x = b.trans_x;
  c2.trans_y = b.trans_y;
  c2.is_additive_component = false;
  if (mode_ == 9) {
    c2.do_warpfield_deformation = b.do_warpfield_deformation;
    c2.field_id = b.field_id;
  }
}

void ParamStream::foreground(TaskBatch& out, size_t idx, bool is_component) {  // DG.cpp:2145-2830
  const bool redraw = (mode_ == 6 || mode_ == 7 || mode_ >= 9);
  const bool thin_modes = (mode_ == 7 || mode_ >= 9);
  {
    ofdg_blueprint b = out.blueprints[idx];
    common_prefix(b, is_component, redraw);
    out.blueprints[idx] = b;
  }
  // A deformed top-level object draws its field when it is realised; components inherit
  // the parent's (DG.cpp:1120-1128, 1157-1169).
  if (mode_ == 9 && !is_component && out.blueprints[idx].do_warpfield_deformation)
    out.blueprints[idx].field_id = next_field();

  const int type = out.blueprints[idx].obj_type;
  const bool type_ok =
      (mode_ == 1 || mode_ == 2) ? (type == OFDG_OBJ_POLYGON)
      : (mode_ == 3)             ? (type == OFDG_OBJ_ELLIPSE)
      : (mode_ == 4 || mode_ == 5 || mode_ == 8) ? (type == OFDG_OBJ_ELLIPSE || type == OFDG_OBJ_POLYGON)
                                 : (type >= OFDG_OBJ_ELLIPSE && type <= OFDG_OBJ_COMPOSITE);
  if (!type_ok) throw std::runtime_error("Bad object type, or not intended in this mode");

  switch (type) {
    case OFDG_OBJ_ELLIPSE: {
      ofdg_blueprint b = out.blueprints[idx];
      ellipse_params(b);
      if (thin_modes && !is_component && eng_[ObjIsExtraThin].trigger()) b.ellipse_scale_x *= 0.05;
      out.blueprints[idx] = b;
      break;
    }
    case OFDG_OBJ_POLYGON: {
      ofdg_blueprint b = out.blueprints[idx];
      polygon_params(out, b, /*curves=*/mode_ >= 4);
      if (thin_modes && !is_component && eng_[ObjIsExtraThin].trigger())
        for (int i = 0; i < b.seg_count; ++i) out.seg_x[b.seg_begin + i] *= 0.05;
      out.blueprints[idx] = b;
      break;
    }
    case OFDG_OBJ_COMPOSITE: {
      if (thin_modes && eng_[ObjIsExtraThin].trigger())
        outline_parts(out, idx);
      else
        composite_parts(out, idx);
      break;
    }
  }
}

void ParamStream::next_task(TaskBatch& out) { next_task_unsynced(out); }

void ParamStream::next_task_unsynced(TaskBatch& out) {  // data_generation_layer.cpp:197-214
  ofdg_blueprint bg = blank_blueprint();
  bg.obj_id = 1;
  bg.obj_type = OFDG_OBJ_POLYGON;
  background(bg);
  out.blueprints.push_back(bg);
  int fg_objs = (int)eng_[NumberOfFgObjects].real();  // DG.cpp:2832-2835: float -> int
  if (fg_override_ > 0) fg_objs = fg_override_;       // stress config only
  for (int obj_idx = 0; obj_idx < fg_objs; ++obj_idx) {
    size_t idx = out.blueprints.size();
    out.blueprints.push_back(blank_blueprint());
    out.blueprints[idx].obj_id = obj_idx + 10;
    foreground(out, idx, false);
  }
  out.task_begin.push_back((int32_t)out.blueprints.size());
  if (augment_) {
    ofdg_augment a{};
    a.enabled = 1;
    std::uniform_real_distribution<double> gain(0.8, 1.2), bright(-20.0, 20.0), contrast(0.7, 1.3), sigma(0.0, 10.0);
    for (int c = 0; c < 3; ++c) a.gain[c] = (float)gain(aug_eng_[0]);
    a.brightness = (float)bright(aug_eng_[1]);
    a.contrast = (float)contrast(aug_eng_[2]);
    a.noise_sigma = (float)sigma(aug_eng_[3]);
    a.noise_seed[0] = aug_eng_[4]();
    a.noise_seed[1] = aug_eng_[4]();
    out.augment.resize(out.task_begin.size() - 2);  // earlier tasks generated without augmentation stay disabled
    out.augment.push_back(a);
  }
  ++tasks_;
}

void ParamStream::set_lookahead_threads(int threads) {
  finish_lookahead();
  pool_.reset();
  if (threads > 0) {
    for (int i = 0; i < kNumSlots; ++i) eng_[i].share();
    pool_.reset(new MiniPool(std::min(threads, 16)));
  }
}

ParamStream::~ParamStream() { finish_lookahead(); }

void ParamStream::finish_lookahead() {
  if (pool_) pool_->wait();
}

// The values of the next n tasks, engine by engine, on the helper threads; busiest engines first. An engine that runs dry
// during the walk simply produces the rest in place, so the estimate only has to be good, not safe.
void ParamStream::start_lookahead(int n) {
  if (!pool_ || !have_rate_ || n < 8) return;
  // keep about two batches' worth of values buffered per engine: the helpers work on the batch after next while the
  // caller walks the next one
  std::vector<std::pair<size_t, int> > want;
  for (int i = 0; i < kNumSlots; ++i) {
    if (rate_[i] <= 0) continue;
    const size_t target = (size_t)(rate_[i] * n * 2.3) + 32, have = eng_[i].buffered();
    if (have + (size_t)(rate_[i] * n * 0.5) < target) want.push_back(std::make_pair(target - have, i));
  }
  std::sort(want.begin(), want.end(), [](const std::pair<size_t, int>& a, const std::pair<size_t, int>& b) { return a.first > b.first; });
  std::vector<std::function<void()> > jobs;
  for (const std::pair<size_t, int>& w : want) {
    Engine* e = &eng_[w.second];
    const size_t cnt = w.first;
    jobs.push_back([e, cnt] { e->fill(cnt); });
  }
  pool_->start(std::move(jobs));
}

void ParamStream::next_tasks(TaskBatch& out, int n) {
  uint64_t before[kNumSlots];
  for (int i = 0; i < kNumSlots; ++i) before[i] = eng_[i].draws;
  for (int i = 0; i < n; ++i) next_task_unsynced(out);
  if (n > 0) {
    for (int i = 0; i < kNumSlots; ++i) {
      const double r = (double)(eng_[i].draws - before[i]) / n;
      rate_[i] = have_rate_ ? 0.5 * rate_[i] + 0.5 * r : r;
    }
    have_rate_ = true;
  }
  start_lookahead(n);  // for the next batch, beside whatever the caller does with this one
}

void ParamStream::skip(uint64_t n_tasks) {
  TaskBatch scratch;
  for (uint64_t i = 0; i < n_tasks; ++i) {
    scratch.clear();
    next_task(scratch);
  }
}

}  // namespace ofdg
