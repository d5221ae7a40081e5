// Host-side scene-parameter stream ("host RNG mode").
//
// Restates, table-driven, what the reference spreads over
//   ObjectParametersGenerator            /root/reference/src/caffe/DataGenerator.cpp:1353-2835
//   RNG::*                               /root/reference/include/caffe/data_generation/SimpleRandom.h:21-142
//   FlyingChairsRandom::*                /root/reference/src/caffe/DataGenerator.cpp:826-922
//   the commission loop of load_batch    /root/reference/src/caffe/layers/data_generation_layer.cpp:197-214
// Every engine owns a private std::mt19937 seeded with its slot index (+ a per-GPU
// offset) and the same libstdc++ distribution class the reference uses, so the
// k-th commissioned task carries the same parameters as the reference's k-th task
// when both are built against the same libstdc++.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "ofdg/scene.h"

namespace ofdg {

enum SlotKind { UINT, UREAL, CHOICE_INT, CHOICE_TYPE, TRIGGER, GAUSS_SQ, GAUSS_3, GAUSS_4, GAUSS_MSR };

struct SlotSpec {
  int kind = UREAL;
  int n_opts = 0;
  int opts[4] = {0, 0, 0, 0};
  double a = 0, b = 0, c = 0, d = 0;  // evaluated in double like the reference's literals, narrowed on use
};

static const int kNumSlots = 45;

// Slot indices (== seeds) in declaration order, DataGenerator.h:524-587.
enum Slot {
  BgTexID = 0, BgInitRot, BgInitTransX, BgInitTransY, BgRotTrigger, BgRot, BgTransX, BgTransY,
  BgScaleTrigger, BgInitScale, BgScale, NumberOfFgObjects, ObjType, ObjTexID, ObjInitTransX,
  ObjInitTransY, ObjTransX, ObjTransY, ObjInitRot, ObjRotTrigger, ObjRot, ObjInitScale,
  ObjScaleTrigger, ObjScale, ObjTexShiftX, ObjTexShiftY, ObjTexRot, ObjTexZoom, ElliObj_ScaleX,
  ElliObj_ScaleY, PolyObj_spokes, PolyObj_dphi, PolyObj_r, PolyObj_ScaleX, PolyObj_ScaleY,
  PolyObj_CurveTrigger, CompObjInitTransX, CompObjInitTransY, CompObiNumberOfComponents,
  ComponentIsAdditive, ComponentOffset, ObjIsExtraThin, ObjDeformsNonrigidly, GenericUniform,
  GenericTrigger
};

// Throws std::runtime_error("BAD MODE") for modes outside 1..13 (DataGenerator.cpp:2003-2005).
void fill_mode_table(int mode, int W, int H, SlotSpec out[kNumSlots]);
const char* slot_name(int slot);

// One seeded engine + its distribution (SimpleRandom.h) + the shaping of
// FlyingChairsRandom. `draws` counts calls, for the stream-bookkeeping tests.
// The k-th value of an engine's stream depends on its seed and on k only (every engine has one fixed distribution), so
// values may be produced ahead of their consumption: prefill() appends the next values of the stream to a buffer (on a
// helper thread, between batches), real() / integer() / trigger() take them from there and fall back to producing in place.
class Engine {
 public:
  Engine() {}
  Engine(const SlotSpec& spec, int seed);
  float real();      // UREAL and the Gaussian family
  int integer();     // UINT, CHOICE_INT, CHOICE_TYPE (returns the chosen option)
  bool trigger();    // TRIGGER
  uint64_t draws = 0;
  // Appends a chunk with the next `count` values of the stream (helper thread). Safe beside the consumer: chunks change
  // hands under the engine's mutex, and an engine that has run dry waits for a fill in progress before it produces in place.
  void fill(size_t count);
  size_t buffered() const { return produced_.load(std::memory_order_relaxed) - consumed_; }  // (consumer thread)
  void share() { shared_ = true; }  // helper threads may call fill() from now on

 private:
  union Value { float f; int i; };
  Value produce();   // the next value of the stream
  Value next() {
    if (pos_ < cur_.size()) { ++consumed_; return cur_[pos_++]; }
    return shared_ ? next_slow() : produce();
  }
  Value next_slow();
  float normal01() { return normal_(mt_); }
  SlotSpec spec_;
  std::mt19937 mt_;
  std::uniform_int_distribution<int> int_;
  std::uniform_real_distribution<double> real_;
  std::normal_distribution<float> normal_;
  float fa_ = 0, fb_ = 0, fc_ = 0, fd_ = 0;
  std::vector<Value> cur_;                 // the chunk being consumed
  size_t pos_ = 0;
  std::deque<std::vector<Value> > ready_;  // filled chunks, in stream order (under mu_)
  std::unique_ptr<std::mutex> mu_;
  std::atomic<size_t> produced_{0};        // values put into chunks so far
  size_t consumed_ = 0;                    // ... and taken out of them
  bool shared_ = false;
  Engine(const Engine&) = delete;

 public:
  Engine(Engine&& o) noexcept { *this = std::move(o); }
  Engine& operator=(Engine&& o) noexcept;
};

// A few persistent helper threads (ParamStream::set_lookahead_threads).
class MiniPool {
 public:
  explicit MiniPool(int threads);
  ~MiniPool();
  void start(std::vector<std::function<void()> > jobs);  // queues them; wait() returns when the last one is done
  void wait();

 private:
  void loop();
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::deque<std::function<void()> > queue_;
  size_t pending_ = 0;
  bool stop_ = false;
};

// Growable owner of the flat arrays behind an ofdg_task_batch.
class TaskBatch {
 public:
  void clear();
  ofdg_task_batch view() const;
  int n_tasks() const { return (int)task_begin.size() - 1; }
  std::vector<int32_t> task_begin{0};
  std::vector<ofdg_blueprint> blueprints;
  std::vector<int32_t> seg_type;
  std::vector<float> seg_x, seg_y;
  std::vector<ofdg_augment> augment;  // empty, or one record per task
};

class ParamStream {
 public:
  // seed_offset shifts all 45 seeds (multi-GPU sharding: 45 * rank); n_fields > 0
  // enables field-id assignment for mode 9 (ids cycle through the injected pool).
  ParamStream(int mode, int W, int H, int seed_offset = 0, int n_fields = 0, int fg_override = 0);
  ~ParamStream();
  // Appends one task (background + foreground objects) to `out`.
  void next_task(TaskBatch& out);
  // Appends n tasks. With look-ahead threads the engines' values for the NEXT batch of this size are produced in parallel (one
  // job per engine, sized from the engines' consumption per task so far) as soon as this one has been drawn -- while the caller
  // flattens and uploads it -- so the sequential walk over the tasks mostly picks finished values up.
  void next_tasks(TaskBatch& out, int n);
  void set_lookahead_threads(int threads);  // 0 (default): every value is produced where it is consumed
  void skip(uint64_t n_tasks);  // fast-forward (checkpoint/resume)
  // Colour/noise augmentation (this repository's own spec, include/ofdg/scene.h): five extra engines seeded
  // 0x40000000 + seed_offset + 0..4, disjoint from every rank's 45 reference seeds. Off by default.
  void enable_augmentation(bool on) { augment_ = on; }
  bool augmentation_enabled() const { return augment_; }
  uint64_t tasks_generated() const { return tasks_; }
  uint64_t field_draws() const { return field_draws_; }  // mode 9: warp-field picks so far (each pool slot serves three)
  uint64_t draws(int slot) const { return eng_[slot].draws; }  // (consumed values; unaffected by the look-ahead)
  int mode() const { return mode_; }

 private:
  void background(ofdg_blueprint& b);
  // Fills blueprint `idx` of `out` (may append component blueprints and segments).
  void foreground(TaskBatch& out, size_t idx, bool is_component);
  void common_prefix(ofdg_blueprint& b, bool is_component, bool redraw_composite);
  void ellipse_params(ofdg_blueprint& b);
  void polygon_params(TaskBatch& out, ofdg_blueprint& b, bool curves);
  void composite_parts(TaskBatch& out, size_t idx);
  void outline_parts(TaskBatch& out, size_t idx);
  int next_field();
  void next_task_unsynced(TaskBatch& out);
  void start_lookahead(int n);
  void finish_lookahead();
  int mode_, W_, H_, n_fields_, fg_override_;
  uint64_t tasks_ = 0, field_draws_ = 0;
  Engine eng_[kNumSlots];
  bool augment_ = false;
  std::mt19937 aug_eng_[5];
  std::unique_ptr<MiniPool> pool_;
  double rate_[kNumSlots] = {};  // values an engine serves per task (from the batches drawn so far)
  bool have_rate_ = false;
};

}  // namespace ofdg
