// Device-side parameter stream + flattening ("Philox mode"); see philox.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "flat_scene.h"
#include "ofdg/scene.h"

namespace ofdg {

constexpr int kPhiloxSlots = 51;        // the 45 engines of a data mode + 5 augmentation engines + the field pick of mode 9
constexpr int kPhiloxMaxObj = 32;       // top-level foreground objects per sample
constexpr int kPhiloxMaxShapes = 8;     // blueprints per object (itself + up to 7 components) = outlines per object
constexpr int kPhiloxMaxBp = 1 + kPhiloxMaxObj * kPhiloxMaxShapes;    // background, then 8 slots per object
constexpr int kPhiloxMaxSeg = kPhiloxMaxObj * kPhiloxMaxShapes * 20;  // polygon segments per sample, 160 per object
constexpr int kPhiloxMaxVerts = 8192;   // fixed-point vertices per object: 512 per (outline, frame)

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
  int n_tex;
  const TexInfo* tex_info;  // device: pool texture sizes
  // mode 9 (n_fields > 0): outlines whose frame-1 masks go through a warp field get a slot of the deformation scratch
  const int* field_reach;   // device [n_fields]: ceil(max |iflow|), how far a warped mask can move
  int* deform_shape;        // [batch * kPhiloxMaxObj * kPhiloxMaxShapes] shape index per slot
  int* deform_field;        //   field id per slot
  int* n_deform;            // device counter (zeroed before the launch)
  int* truncated;           // mapped host int (may be null): raised when a scene does not fit the fixed strides below and loses objects / outlines
  // blueprints in the ABI layout, fixed strides per sample (downloadable for inspection / the oracle)
  ofdg_blueprint* bp;   // [batch][kPhiloxMaxBp]: background, then kPhiloxMaxShapes slots per object
  int32_t* seg_type;    // [batch][kPhiloxMaxSeg]: 160 per object
  float* seg_x;
  float* seg_y;
  int* obj_nbp;         // [batch][kPhiloxMaxObj] blueprints / segments object k actually uses
  int* obj_nseg;
  int* n_top;           // [batch] foreground objects
  // flattened scene, fixed strides
  FlatSample* samples;  // [batch]
  FlatObject* objects;  // [batch][kPhiloxMaxObj]
  FlatShape* shapes;    // [batch][kPhiloxMaxObj][kPhiloxMaxShapes]
  FlatVertex* verts;    // [batch][kPhiloxMaxObj][kPhiloxMaxVerts]
};

void philox_upload_circle(const double* cos100, const double* sin100);
int launch_philox(const PhiloxArgs& a, cudaStream_t s);

}  // namespace ofdg
