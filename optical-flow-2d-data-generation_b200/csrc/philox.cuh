// Device-side parameter stream + flattening ("Philox mode"); see philox.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "flat_scene.h"
#include "ofdg/scene.h"

namespace ofdg {

constexpr int kPhiloxSlots = 50;        // the 45 engines of a data mode + 5 augmentation engines
constexpr int kPhiloxMaxObj = 32;       // top-level foreground objects per sample
constexpr int kPhiloxMaxBp = 1 + kPhiloxMaxObj * 8;   // background + objects + up to 7 components each
constexpr int kPhiloxMaxSeg = kPhiloxMaxObj * 8 * 20; // polygon segments per sample
constexpr int kPhiloxMaxShapes = 8;     // outlines per object
constexpr int kPhiloxMaxVerts = 4096;   // fixed-point vertices per object (all outlines, both frames)

struct PhiloxSlot {  // one row of a mode table, narrowed to float like the reference's constructors do
  int kind;          // ofdg::SlotKind
  int n_opts;
  int opts[4];
  int ia, ib;        // UINT bounds
  float a, b, c, d;
};

struct PhiloxArgs {
  const PhiloxSlot* slots;  // device, kPhiloxSlots rows (only the first 45 are table rows)
  int mode, W, H;
  uint64_t seed, first_sample;
  int batch, n_fields, fg_override, augment;
  int n_tex, tex_w, tex_h;
  // blueprints in the ABI layout, fixed strides per sample (downloadable for inspection / the oracle)
  ofdg_blueprint* bp;   // [batch][kPhiloxMaxBp]
  int* bp_count;        // [batch]
  int32_t* seg_type;    // [batch][kPhiloxMaxSeg]
  float* seg_x;
  float* seg_y;
  int* seg_count;       // [batch]
  int* top_index;       // [batch][kPhiloxMaxObj] blueprint index (within the sample) of the k-th object
  int* n_top;           // [batch]
  // flattened scene, fixed strides
  FlatSample* samples;  // [batch]
  FlatObject* objects;  // [batch][kPhiloxMaxObj]
  FlatShape* shapes;    // [batch][kPhiloxMaxObj][kPhiloxMaxShapes]
  FlatVertex* verts;    // [batch][kPhiloxMaxObj][kPhiloxMaxVerts]
};

void philox_upload_circle(const double* cos100, const double* sin100);
int launch_philox(const PhiloxArgs& a, cudaStream_t s);

}  // namespace ofdg
