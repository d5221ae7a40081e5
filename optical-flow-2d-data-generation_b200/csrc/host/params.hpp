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
#include <cstdint>
#include <random>
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
class Engine {
 public:
  Engine() {}
  Engine(const SlotSpec& spec, int seed);
  float real();      // UREAL and the Gaussian family
  int integer();     // UINT, CHOICE_INT, CHOICE_TYPE (returns the chosen option)
  bool trigger();    // TRIGGER
  uint64_t draws = 0;

 private:
  float normal01() { return normal_(mt_); }
  SlotSpec spec_;
  std::mt19937 mt_;
  std::uniform_int_distribution<int> int_;
  std::uniform_real_distribution<double> real_;
  std::normal_distribution<float> normal_;
  float fa_ = 0, fb_ = 0, fc_ = 0, fd_ = 0;
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
  // Appends one task (background + foreground objects) to `out`.
  void next_task(TaskBatch& out);
  void skip(uint64_t n_tasks);  // fast-forward (checkpoint/resume)
  // Colour/noise augmentation (this repository's own spec, include/ofdg/scene.h): five extra engines seeded
  // seed_offset + 45..49, so the reference's 45 streams are untouched. Off by default.
  void enable_augmentation(bool on) { augment_ = on; }
  bool augmentation_enabled() const { return augment_; }
  uint64_t tasks_generated() const { return tasks_; }
  uint64_t field_draws() const { return field_draws_; }  // mode 9: warp-field picks so far (each pool slot serves three)
  uint64_t draws(int slot) const { return eng_[slot].draws; }
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
  int mode_, W_, H_, n_fields_, fg_override_;
  uint64_t tasks_ = 0, field_draws_ = 0;
  Engine eng_[kNumSlots];
  bool augment_ = false;
  std::mt19937 aug_eng_[5];
};

}  // namespace ofdg
