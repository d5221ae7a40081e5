// Flattened scene records: what the host geometry stage (host/flatten.cpp) uploads and the
// sm_100a kernels (render.cu) consume. Plain data, identical layout on host and device.
//
// Matrix convention everywhere: double m[6] = {sx, shy, shx, sy, tx, ty} (AGG order),
//   x' = x*sx + y*shx + tx,   y' = x*shy + y*sy + ty.
#pragma once
#include <stdint.h>

#include "ofdg/scene.h"

namespace ofdg {

// One closed outline (a simple object, or one component of a composite), both frames.
// Vertices are AGG 24.8 fixed-point pairs; edge i runs vertex i -> vertex (i+1) % count.
struct FlatShape {
  int32_t vbegin[2];    // [frame] first vertex in FlatBatch::verts
  int32_t vcount[2];
  int32_t bbox[2][4];   // [frame] {x0, y0, x1, y1} inclusive pixel range that can hold cells
  int32_t additive;     // composite op: 1 = add, 0 = subtract (DataGenerator.cpp:602-642)
  int32_t deform;       // mode 9: slot of this outline's warped frame-1 masks in the deformation scratch, else -1
  int32_t raw1[4];      // mode 9: frame-1 box before it was widened by the field's reach (what the pre-pass rasterises)
};

// One top-level foreground object, in z-order within its sample.
struct FlatObject {
  double tex_inv[6];    // inverse of the motion: frame-1 texture lookup (DataGenerator.cpp:203-207)
  double motion[6];     // forward motion incl. background motion: flow (DataGenerator.cpp:388-401)
  int32_t bbox[2][4];   // [frame] union of the shape boxes
  int32_t shape_begin, shape_count;
  int32_t tex;          // pool slot (tex_id % pool size, DataGenerator.cpp:158-161)
  int32_t obj_id;       // 10 + k
  int32_t composite;    // masks are built with the ADD/SUB rules even for one component
  int32_t field;        // mode 9 field-pool slot or -1
  int32_t pad[2];
};

// One texture of the HBM pool (RGBX8 pixels, textures of any size back to back; TextureCollection,
// DataGenerator.cpp:117-161). Foreground objects see Texture::getRandomizedCrop(W, H) with its default
// arguments (DataGenerator.cpp:87-109): the centre W x H window of a texture that is at least W x H,
// else the whole texture resized to W x H -- that resized copy is made once at upload and lives in the pool too.
struct TexInfo {
  uint64_t off;        // first pixel of the raw texture
  uint64_t fg_base;    // first pixel of the foreground view (W x H): raw + centre-crop origin, or the resized copy
  int32_t w, h;        // raw size
  int32_t fg_pitch;    // row pitch of the foreground view in pixels
  int32_t pad;
};

// Background texture preparation = Texture::getRandomizedCrop(2W, 2H, rot, zoom, sx, sy)
// (DataGenerator.cpp:87-109) restated as closed-form per-pixel parameters.
struct BgPrep {
  int32_t tex;                       // pool slot
  int32_t shift_x, shift_y;          // get_shift(dx, dy, 0, 0, mirror)
  int32_t rot_identity;              // rotate() degenerates to a copy
  float ca, sa;                      // cos / sin of the (degree-valued) angle, as float
  float w2, h2, rw2, rh2;            // source / rotated image centres
  int32_t rw, rh;                    // rotated image size (bounding box)
  int32_t crop_x0, crop_y0;          // crop origin in the rotated image
  int32_t crop_w, crop_h;            // crop size (then resized to 2W x 2H)
  int32_t need[4];                   // {x0, y0, x1, y1}: part of the prepared texture the renderer reads
  int32_t general;                   // the crop is more than 1.3x the prepared size on an axis (textures smaller than 2W x 2H
  int32_t pad;                       //   skip the crop, DataGenerator.cpp:103-107): sub-tiled path with unbounded tap counts
};

struct FlatSample {
  int32_t obj_begin, obj_count;      // foreground objects in FlatBatch::objects
  int32_t bg_field;                  // mode 9 field-pool slot or -1
  int32_t pad;
  double bg_tex_inv[6];              // inverse of I^-1 * M * I on the 2W x 2H canvas (DataGenerator.cpp:676-677)
  double bg_motion[6];               // M alone; the flow applies I^-1, M, I in turn (DataGenerator.cpp:692-712)
  double bg_motion_inv[6];           // M^-1 (m_motion_inv): backward flow, getPointFlow(inverse = true)
  BgPrep prep;
  ofdg_augment aug;                  // colour/noise augmentation of this sample (enabled == 0: none)
};

struct FlatVertex {
  int32_t x, y;
};

}  // namespace ofdg
