// Host geometry stage: blueprints -> fixed-point outlines + per-object matrices.
// Everything here is IEEE-double arithmetic in the order AGG 2.4 performs it, because the
// results are rounded to 24.8 fixed point and must land on the same sub-pixel as the
// reference's. Replaces, per object,
//   setIntrinsicTransform / setMotion / addBackgroundMotion   DataGenerator.cpp:302-335
//   setEllipse + conv_transform<ellipse>                       DataGenerator.cpp:459-476
//   path_storage + conv_transform + conv_curve                 DataGenerator.cpp:491-531
//   RealizeObjectBlueprint's object construction               DataGenerator.cpp:1065-1173
// and, per sample, the parameter side of Texture::getRandomizedCrop (DataGenerator.cpp:87-109).
#pragma once
#include <vector>

#include "../flat_scene.h"
#include "ofdg/scene.h"

namespace ofdg {

struct FlatBatch {
  std::vector<FlatSample> samples;
  std::vector<FlatObject> objects;
  std::vector<FlatShape> shapes;
  std::vector<FlatVertex> verts;
  // mode 9: per deformation-scratch slot, the outline it belongs to and the field that warps it
  std::vector<int32_t> deform_shape, deform_field;
  void clear() { samples.clear(); objects.clear(); shapes.clear(); verts.clear(); deform_shape.clear(); deform_field.clear(); }
};

struct FlattenConfig {
  int W = 512, H = 384;      // output size (DGEN_WIDTH / DGEN_HEIGHT)
  const TexInfo* tex_info = nullptr;  // pool textures (sizes), n_tex entries
  int n_tex = 0;             // pool size
  int mode = 1;              // only mode 9 attaches warp fields
  int n_fields = 0;          // injected field pool
  const int* field_reach = nullptr;  // per field: ceil(max |iflow|) over finite entries (how far a warped mask can move)
};

// Appends the flattened form of every task in `tb` to `out`.
// Throws std::runtime_error on descriptors the reference would also reject.
void flatten(const ofdg_task_batch& tb, const FlattenConfig& cfg, FlatBatch& out);

// Exposed for the known-answer tests.
void flatten_ellipse(double rx, double ry, const double m[6], std::vector<FlatVertex>& out);
void flatten_polygon(const int32_t* seg_type, const float* seg_x, const float* seg_y, int n,
                     const double m[6], std::vector<FlatVertex>& out);

}  // namespace ofdg
