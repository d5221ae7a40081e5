/*
 * ofdg/scene.h -- plain-old-data scene records that cross the C ABI.
 *
 * One "task" is one training sample: one background blueprint (obj_id 1) plus
 * 16..23 foreground blueprints (obj_id 10+k), exactly the record the reference
 * passes from its layer to its generator:
 *   ObjectBlueprint  /root/reference/include/caffe/data_generation/DataGenerator.h:388-421
 *   TaskBucket       /root/reference/include/caffe/data_generation/DataGenerator.h:423-437
 * The reference stores polygon segments and composite components in std::vectors
 * hanging off each blueprint; here they are index ranges into flat arrays of a
 * task batch so the whole batch is five contiguous arrays.
 */
#ifndef OFDG_SCENE_H_
#define OFDG_SCENE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ObjType_t, DataGenerator.h:369-374 */
enum { OFDG_OBJ_DUMMY = 0, OFDG_OBJ_ELLIPSE = 1, OFDG_OBJ_POLYGON = 2, OFDG_OBJ_COMPOSITE = 3 };
/* PolySegmentType_t, DataGenerator.h:377-381 */
enum { OFDG_SEG_DUMMY = 0, OFDG_SEG_LINE = 1, OFDG_SEG_CURVE3 = 3 };

/* One object (or composite component) description. 1:1 with the reference's
 * ObjectBlueprint fields; the fields the reference leaves uninitialised
 * are zero here. 100 bytes. */
typedef struct ofdg_blueprint {
  int32_t obj_id;                 /* 1 = background, 10+k = k-th foreground object, 0 = component */
  int32_t obj_type;               /* OFDG_OBJ_* */
  float   init_rot;               /* intrinsic transform I = R(init_rot) * T(init_trans) */
  float   init_scale;             /* never used by the reference */
  float   init_trans_x, init_trans_y;
  float   rot, scale;             /* motion M = R(rot) * S(scale) * T(trans) */
  float   trans_x, trans_y;
  int32_t tex_id;                 /* raw random index, taken modulo the pool size */
  float   tex_rot, tex_scale;     /* background only: texture rotation ("degrees") / zoom */
  int32_t tex_shift_x, tex_shift_y;
  float   ellipse_scale_x, ellipse_scale_y;
  int32_t seg_begin, seg_count;   /* polygon: range in the batch's seg_type/seg_x/seg_y arrays */
  int32_t comp_begin, comp_count; /* composite: range of component blueprints in the batch's blueprint array */
  int32_t parent;                 /* index (in the batch's blueprint array) of the composite this is a component of, else -1 */
  int32_t is_additive_component;
  int32_t do_warpfield_deformation;
  int32_t field_id;               /* mode 9: index into the injected (flow, iflow) field pool, -1 = none */
} ofdg_blueprint;

/* Per-sample colour / noise augmentation. NOT part of the reference (it has no
 * augmentation anywhere); this is this repository's own, deliberately simple specification so that
 * the oracle reproduces it bit for bit: for frame f, channel c, pixel p with 8-bit value v,
 *   y = contrast * (gain[c] * v - 127.5f) + 127.5f + brightness + noise_sigma * n(f, c, p)
 *   out = min(max(y, 0), 255)                       (float32, one rounding per operation, no FMA)
 * n = ((sum of the four bytes of word c of Philox4x32-10(key = noise_seed, counter = {p, f, 0, 0}))
 *      - 510) * (1 / 147.80f)                       (Irwin-Hall of four bytes, approximately N(0,1), integer-exact;
 *                                                     one Philox call serves the three channels of a pixel: ofdg/augment.h)
 * The flow is not touched. enabled == 0 leaves the sample exactly as the reference would produce it. */
typedef struct ofdg_augment {
  int32_t  enabled;
  float    gain[3];
  float    brightness;
  float    contrast;
  float    noise_sigma;
  uint32_t noise_seed[2];
} ofdg_augment;

/* A batch of tasks as flat arrays (all owned by whoever built the batch).
 * Task t owns blueprints [task_begin[t], task_begin[t+1]); the first one is the
 * background; the top-level foreground objects are those with parent == -1, in
 * array order (= ascending obj_id = z-order, later on top). */
typedef struct ofdg_task_batch {
  int32_t               n_tasks;
  int32_t               n_blueprints;
  int32_t               n_segments;
  const int32_t*        task_begin;    /* n_tasks + 1 */
  const ofdg_blueprint* blueprints;    /* n_blueprints */
  const int32_t*        seg_type;      /* n_segments, OFDG_SEG_* */
  const float*          seg_x;         /* n_segments */
  const float*          seg_y;         /* n_segments */
  const ofdg_augment*   augment;       /* n_tasks records, or NULL (= no augmentation) */
} ofdg_task_batch;

#ifdef __cplusplus
}
#endif
#endif /* OFDG_SCENE_H_ */
