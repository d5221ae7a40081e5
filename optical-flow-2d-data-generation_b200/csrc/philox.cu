// Device-side ("production") scene generation: parameters drawn with counter-based Philox4x32-10 and
// geometry flattened on the GPU, so that no host work is left on the path (the host-RNG
// stream of host/params.cpp is the parity mode and stays the reference-comparable one).
//
//   philox_params_kernel    one thread per (sample, object): ObjectParametersGenerator's logic
//                           (/root/reference/src/caffe/DataGenerator.cpp:2105-2835, same mode tables, same
//                           branch structure) with every engine replaced by Philox keyed on
//                           (seed, sample index, object, slot, draw number) -> blueprints in the ABI's POD layout
//   philox_flatten_kernel   one thread per (sample, object, outline, frame): host/flatten.cpp on the device
//
// A sample is a pure function of (mode, seed, sample index): any GPU reproduces any sample. The streams
// are statistically, not bitwise, equal to the host mode's (different engine, device libm).
#include "philox.cuh"

#include "host/affine.hpp"
#include "ofdg/augment.h"

namespace ofdg {

namespace {

__constant__ double c_circle_cos[100];
__constant__ double c_circle_sin[100];

// ---- Philox-backed engines ---------------------------------------------------------------------------
struct Rng {
  uint32_t k0, k1, s0, s1, obj;
  unsigned char ctr[kPhiloxSlots];
  __device__ void init(uint64_t seed, uint64_t sample, int object) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); s0 = (uint32_t)sample; s1 = (uint32_t)(sample >> 32);
    obj = (uint32_t)object << 8;
    for (int i = 0; i < kPhiloxSlots; ++i) ctr[i] = 0;
  }
  __device__ void raw(int slot, uint32_t& a, uint32_t& b) { raw_at(slot, 0, a, b); ++ctr[slot]; }
  // the draw `ahead` positions further down the slot's stream, without consuming anything (see Gen::lanes)
  __device__ void raw_at(int slot, int ahead, uint32_t& a, uint32_t& b) const {
    uint32_t x0 = s0, x1 = s1, x2 = (uint32_t)slot | obj, x3 = (uint32_t)(unsigned char)(ctr[slot] + ahead);
    uint32_t ka = k0, kb = k1;
    for (int r = 0; r < 10; ++r) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
      const uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ ka, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ x3 ^ kb, y3 = (uint32_t)p0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    a = x0; b = x1;
  }
  __device__ float unit(int slot, int ahead = -1) {  // [0, 1); ahead >= 0: that draw of the stream, nothing consumed
    uint32_t a, b;
    if (ahead < 0) raw(slot, a, b); else raw_at(slot, ahead, a, b);
    return (float)(a >> 8) * (1.0f / 16777216.0f);
  }
  __device__ float normal(int slot, int ahead = -1) {  // Box-Muller
    uint32_t a, b;
    if (ahead < 0) raw(slot, a, b); else raw_at(slot, ahead, a, b);
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
  }
  __device__ void skip(int slot, int n) { ctr[slot] = (unsigned char)(ctr[slot] + n); }
};

struct Gen {
  const PhiloxSlot* slots;
  Rng rng;
  int mode;
  __device__ float base_gauss(float a, float b, float input, float normalize) {  // DG.cpp:828-831
    const float sample = input * ((b + a) / 2.f - a) / normalize + (b + a) / 2.f;
    return (a <= sample && sample <= b) ? sample : (b + a) / 2.f;
  }
  // ahead >= 0 (here and in trigger): the value the ahead-th next call would return, without consuming a draw. Every draw is
  // a pure function of (seed, sample, object, slot, draw number), so the lanes of the warp can each take one of a run of
  // consecutive draws of a slot (a polygon's spokes) instead of one lane walking the run; rng.skip then consumes the run.
  __device__ float real(int slot, int ahead = -1) {
    const PhiloxSlot& s = slots[slot];
    switch (s.kind) {
      case 1: return s.a + (s.b - s.a) * rng.unit(slot, ahead);                                     // UREAL
      case 5: { float t = rng.normal(slot, ahead); t = t > 0 ? t * t : -(t * t); return base_gauss(s.a, s.b, t, 6); }   // GAUSS_SQ
      case 6: { float t = rng.normal(slot, ahead); return base_gauss(s.a, s.b, t * t * t, 10); }      // GAUSS_3
      case 7: { float t = rng.normal(slot, ahead); const float q = t * t * t * t; return base_gauss(s.a, s.b, t > 0 ? q : -q, 15); }  // GAUSS_4
      default: { float t = rng.normal(slot, ahead) * s.d + s.c; return (s.a <= t && t <= s.b) ? t : s.c; }  // GAUSS_MSR
    }
  }
  __device__ int integer(int slot) {
    const PhiloxSlot& s = slots[slot];
    uint32_t a, b;
    rng.raw(slot, a, b);
    if (s.kind == 0) {  // UINT in [a, b]
      const uint64_t range = (uint64_t)((long long)s.ib - (long long)s.ia) + 1ull;
      return s.ia + (int)(((uint64_t)a * range) >> 32);
    }
    return s.opts[(int)(((uint64_t)a * (uint64_t)s.n_opts) >> 32)];  // CHOICE_*
  }
  __device__ bool trigger(int slot, int ahead = -1) {
    const PhiloxSlot& s = slots[slot];
    return s.a + (s.b - s.a) * rng.unit(slot, ahead) < s.c;
  }
};

enum {  // slot indices, DataGenerator.h:524-587
  BgTexID = 0, BgInitRot, BgInitTransX, BgInitTransY, BgRotTrigger, BgRot, BgTransX, BgTransY, BgScaleTrigger, BgInitScale,
  BgScale, NumberOfFgObjects, ObjType, ObjTexID, ObjInitTransX, ObjInitTransY, ObjTransX, ObjTransY, ObjInitRot, ObjRotTrigger,
  ObjRot, ObjInitScale, ObjScaleTrigger, ObjScale, ObjTexShiftX, ObjTexShiftY, ObjTexRot, ObjTexZoom, ElliObj_ScaleX,
  ElliObj_ScaleY, PolyObj_spokes, PolyObj_dphi, PolyObj_r, PolyObj_ScaleX, PolyObj_ScaleY, PolyObj_CurveTrigger,
  CompObjInitTransX, CompObjInitTransY, CompObiNumberOfComponents, ComponentIsAdditive, ComponentOffset, ObjIsExtraThin,
  ObjDeformsNonrigidly, GenericUniform, GenericTrigger, AugGain, AugBrightness, AugContrast, AugSigma, AugSeed, FieldPick
};

struct SampleOut {  // this sample's slice of the blueprint arrays
  ofdg_blueprint* bp;
  int32_t* seg_type;
  float* seg_x;
  float* seg_y;
  int nbp, nseg;
  int bp_base, seg_base;  // absolute offsets (the batch is one ofdg_task_batch)
};

__device__ ofdg_blueprint blank_bp() {
  ofdg_blueprint b;
  memset(&b, 0, sizeof(b));
  b.parent = -1;
  b.field_id = -1;
  return b;
}

// Every lane of the role's warp holds the same blueprint; lane 0 writes it to the staging area, and the warp is
// synchronised so that any lane may read it back.
__device__ void put_bp(SampleOut& o, int idx, const ofdg_blueprint& b) {
  __syncwarp();  // (no lane is still reading the slot's previous content)
  if ((threadIdx.x & 31) == 0) o.bp[idx] = b;
  __syncwarp();
}

__device__ void gen_polygon(Gen& g, SampleOut& o, ofdg_blueprint& b, bool curves) {
  b.seg_begin = o.seg_base + o.nseg;
  float* sx = o.seg_x + o.nseg;
  float* sy = o.seg_y + o.nseg;
  int32_t* st = o.seg_type + o.nseg;
  if (g.mode == 1) {  // DG.cpp:2163-2183
    const float radius = g.real(PolyObj_r);
    const float xs = radius * g.real(PolyObj_ScaleX), ys = radius * g.real(PolyObj_ScaleY);
    if ((threadIdx.x & 31) == 0) {
      sx[0] = xs; sx[1] = xs; sx[2] = -xs; sx[3] = -xs;
      sy[0] = -ys; sy[1] = ys; sy[2] = ys; sy[3] = -ys;
      st[0] = OFDG_SEG_DUMMY; st[1] = st[2] = st[3] = OFDG_SEG_LINE;
    }
    __syncwarp();
    b.seg_count = 4;
    o.nseg += 4;
    return;
  }
  // The whole warp runs this code with identical state; where the reference loops over the spokes, lane i takes spoke i
  // (its draws are the i-th next ones of their slots) and writes its own segment.
  const int lane = threadIdx.x & 31;
  const int spokes = g.integer(PolyObj_spokes);  // DG.cpp:2469-2495 (at most 20: the mode tables' choices)
  float px = 0.f, py = 0.f;
  for (int i = lane; i < spokes; i += 32) {
    const float phi = (float)((i * 360. / spokes + g.real(PolyObj_dphi, i)) * 3.14159265358979323846 / 180.);
    const float r = g.real(PolyObj_r, i);
    px = r * cosf(phi);  // scaled below, once ScaleX / ScaleY are drawn (same draw order as the reference)
    py = r * sinf(phi);
  }
  g.rng.skip(PolyObj_dphi, spokes);
  g.rng.skip(PolyObj_r, spokes);
  const float xscale = g.real(PolyObj_ScaleX), yscale = g.real(PolyObj_ScaleY);
  // the curve triggers: the n-th one asked for is the n-th next draw, whichever spoke asks; lane n draws it, every lane replays the walk
  const unsigned fired = __ballot_sync(0xffffffffu, curves && lane < spokes && g.trigger(PolyObj_CurveTrigger, lane));
  int asked = 0, mine = OFDG_SEG_LINE;
  for (int i = 1; i < spokes; ++i) {
    if (curves && (i < spokes - 1) && ((fired >> asked++) & 1u)) {
      if (lane == i) mine = OFDG_SEG_CURVE3;
      if (lane == i + 1) mine = OFDG_SEG_DUMMY;
      ++i;
    }
  }
  g.rng.skip(PolyObj_CurveTrigger, asked);
  if (lane == 0) mine = OFDG_SEG_DUMMY;
  if (lane < spokes) { sx[lane] = px * xscale; sy[lane] = py * yscale; st[lane] = mine; }
  __syncwarp();
  b.seg_count = spokes;
  o.nseg += spokes;
}

__device__ void shrink(SampleOut& o, ofdg_blueprint& c, float f) {  // (whole warp: lanes over the segments)
  if (c.obj_type == OFDG_OBJ_ELLIPSE) { c.ellipse_scale_x *= f; c.ellipse_scale_y *= f; }
  else {
    for (int i = threadIdx.x & 31; i < c.seg_count; i += 32) { o.seg_x[c.seg_begin - o.seg_base + i] *= f; o.seg_y[c.seg_begin - o.seg_base + i] *= f; }
    __syncwarp();
  }
}

// generateForegroundObject (DG.cpp:2145-2830); components are never composite, so one level of nesting suffices
__device__ void gen_simple(Gen& g, SampleOut& o, int idx, bool is_component, int n_fields, int& field_draws) {
  ofdg_blueprint b = o.bp[idx];
  const bool redraw = (g.mode == 6 || g.mode == 7 || g.mode >= 9), thin_modes = (g.mode == 7 || g.mode >= 9);
  do { b.obj_type = g.integer(ObjType); } while (redraw && is_component && b.obj_type == OFDG_OBJ_COMPOSITE);
  b.init_rot = g.real(ObjInitRot);
  b.init_trans_x = g.real(ObjInitTransX);
  b.init_trans_y = g.real(ObjInitTransY);
  b.rot = g.trigger(ObjRotTrigger) ? g.real(ObjRot) : 0.f;
  b.scale = g.trigger(ObjScaleTrigger) ? g.real(ObjScale) : 1.f;
  b.trans_x = g.real(ObjTransX);
  b.trans_y = g.real(ObjTransY);
  b.tex_id = g.integer(ObjTexID);
  if (g.mode == 9) {
    b.do_warpfield_deformation = g.trigger(ObjDeformsNonrigidly);
    // the host stream walks the injected pool in commission order; a sample of the device stream must not depend on
    // the samples before it, so the pick is one more counter-based draw
    if (!is_component && b.do_warpfield_deformation && n_fields > 0) {
      uint32_t r0, r1;
      g.rng.raw(FieldPick, r0, r1);
      b.field_id = (int)(r0 % (uint32_t)n_fields);
      ++field_draws;
    }
  }
  if (b.obj_type == OFDG_OBJ_ELLIPSE) {
    b.ellipse_scale_x = g.real(ElliObj_ScaleX) * 50;
    b.ellipse_scale_y = g.real(ElliObj_ScaleY) * 50;
    if (thin_modes && !is_component && g.trigger(ObjIsExtraThin)) b.ellipse_scale_x *= 0.05f;
  } else if (b.obj_type == OFDG_OBJ_POLYGON) {
    gen_polygon(g, o, b, g.mode >= 4);
    if (thin_modes && !is_component && g.trigger(ObjIsExtraThin)) {
      for (int i = threadIdx.x & 31; i < b.seg_count; i += 32) o.seg_x[b.seg_begin - o.seg_base + i] *= 0.05f;
      __syncwarp();
    }
  }
  put_bp(o, idx, b);
}

__device__ void copy_placement(ofdg_blueprint& c, const ofdg_blueprint& b) {
  c.init_rot = b.init_rot; c.init_trans_x = b.init_trans_x; c.init_trans_y = b.init_trans_y;
  c.rot = b.rot; c.scale = b.scale; c.trans_x = b.trans_x; c.trans_y = b.trans_y;
}

__device__ void gen_object(Gen& g, SampleOut& o, int idx, int n_fields, int& field_draws) {
  gen_simple(g, o, idx, false, n_fields, field_draws);
  if (o.bp[idx].obj_type != OFDG_OBJ_COMPOSITE) return;
  const bool thin_modes = (g.mode == 7 || g.mode >= 9);
  ofdg_blueprint b = o.bp[idx];
  b.comp_begin = o.bp_base + o.nbp;
  if (thin_modes && g.trigger(ObjIsExtraThin)) {  // "outline": a shape minus a slightly smaller copy (DG.cpp:2504-2547)
    const int i1 = o.nbp++;
    {
      ofdg_blueprint nb = blank_bp();
      nb.obj_type = OFDG_OBJ_COMPOSITE;
      put_bp(o, i1, nb);
    }
    gen_simple(g, o, i1, true, n_fields, field_draws);
    ofdg_blueprint c1 = o.bp[i1];
    c1.parent = o.bp_base + idx;
    copy_placement(c1, b);
    c1.is_additive_component = 1;
    c1.do_warpfield_deformation = b.do_warpfield_deformation; c1.field_id = b.field_id;
    put_bp(o, i1, c1);
    const int i2 = o.nbp++;
    ofdg_blueprint c2 = c1;
    if (c1.obj_type == OFDG_OBJ_POLYGON) {
      c2.seg_begin = o.seg_base + o.nseg;
      for (int i = threadIdx.x & 31; i < c1.seg_count; i += 32) {
        o.seg_type[o.nseg + i] = o.seg_type[c1.seg_begin - o.seg_base + i];
        o.seg_x[o.nseg + i] = o.seg_x[c1.seg_begin - o.seg_base + i];
        o.seg_y[o.nseg + i] = o.seg_y[c1.seg_begin - o.seg_base + i];
      }
      __syncwarp();
      o.nseg += c1.seg_count;
    }
    if (c1.obj_type == OFDG_OBJ_ELLIPSE) {
      if (g.trigger(GenericTrigger)) {
        c2.init_trans_x = b.init_trans_x + g.real(CompObjInitTransX);
        c2.init_trans_y = b.init_trans_y + g.real(CompObjInitTransY);
      } else {
        c2.ellipse_scale_x *= 0.9f; c2.ellipse_scale_y *= 0.9f;
      }
    } else {
      shrink(o, c2, 0.9f);
    }
    c2.is_additive_component = 0;
    put_bp(o, i2, c2);
    b.comp_count = 2;
  } else {  // DG.cpp:2549-2591
    const int parts = g.integer(CompObiNumberOfComponents);
    for (int part = 0; part < parts; ++part) {
      const int ci = o.nbp++;
      {
        ofdg_blueprint nb = blank_bp();
        nb.obj_type = OFDG_OBJ_COMPOSITE;
        put_bp(o, ci, nb);
      }
      gen_simple(g, o, ci, true, n_fields, field_draws);
      ofdg_blueprint c = o.bp[ci];
      c.parent = o.bp_base + idx;
      copy_placement(c, b);
      if (part == 0) {
        c.is_additive_component = 1;
      } else {
        c.init_rot = g.real(ObjInitRot);
        c.init_trans_x += g.real(ComponentOffset);
        c.init_trans_y += g.real(ComponentOffset);
        shrink(o, c, 0.2f);
        c.is_additive_component = g.trigger(ComponentIsAdditive) ? 1 : 0;
      }
      c.do_warpfield_deformation = b.do_warpfield_deformation; c.field_id = b.field_id;
      put_bp(o, ci, c);
    }
    b.comp_count = parts;
  }
  put_bp(o, idx, b);
}

// One WARP per (sample, role): the roles of a sample (its objects, its background) run through very different branches and
// loop counts, so as threads of one warp they would execute one after the other. All 32 lanes of the role's warp run its code
// with identical state (no divergence, nothing wasted that lane 0 alone would not idle away), and split the loops over a
// polygon's spokes / segments among them: the longest role (a composite of seven 20-spoke polygons) drops from ~500 draws in a
// row to ~100. An object's
// blueprints and polygon segments are read, scaled and rewritten several times while they are drawn (components copy and
// shrink their parent's outline, thin objects rescale theirs): they live in shared memory until the object is complete and
// leave in one coalesced copy by the whole warp -- as global read-modify-write chains they were most of the kernel's 169 us,
// during which its blocks (17 warps at 56 registers for one working lane each) held half an SM's registers away from the
// render kernels of the batch before.
#ifndef OFDG_PARAMS_MIN_BLOCKS
#define OFDG_PARAMS_MIN_BLOCKS 5  // 3 / 4 / 5 blocks per SM (56 / 40 / 32 registers, spilling): production mode 142.5k / 143.5k / 146.5k samples/s -- what the kernel costs the render beside it is the registers its blocks hold
#endif
constexpr int kParamWarps = 12;                                          // roles per block
constexpr int kParamBlocks = (kPhiloxMaxObj + 1 + kParamWarps - 1) / kParamWarps;  // blocks per sample
constexpr int kSegPerObj = kPhiloxMaxShapes * 20;
struct ParamStage {
  ofdg_blueprint bp[kPhiloxMaxShapes];
  int32_t seg_type[kSegPerObj];
  float seg_x[kSegPerObj];
  float seg_y[kSegPerObj];
};
__global__ void __launch_bounds__(32 * kParamWarps, OFDG_PARAMS_MIN_BLOCKS) philox_params_kernel(PhiloxArgs a) {
  __shared__ ParamStage s_stage[kParamWarps];
  __shared__ PhiloxSlot s_slots[kPhiloxSlots];
  for (int i = threadIdx.x; i < (int)(kPhiloxSlots * sizeof(PhiloxSlot) / 4); i += blockDim.x)
    reinterpret_cast<int*>(s_slots)[i] = reinterpret_cast<const int*>(a.slots)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int s = blockIdx.x, k = blockIdx.y * kParamWarps + w;  // sample, role: object k, or kPhiloxMaxObj = background and sample-level draws
  if (k > kPhiloxMaxObj) return;
  ParamStage& st = s_stage[w];
  const int bp_base = s * kPhiloxMaxBp + 1 + k * kPhiloxMaxShapes, seg_base = s * kPhiloxMaxSeg + k * kSegPerObj;
  int nbp = 0, nseg = 0;
  {  // every lane runs the role's code with identical state (same draws, same branches); the lanes only differ inside the loops over a polygon's segments
    Gen g;
    g.slots = s_slots;
    g.mode = a.mode;
    // the number of objects is a sample-level draw every role of the sample repeats
    g.rng.init(a.seed, a.first_sample + (uint64_t)s, kPhiloxMaxObj);
    int fg = a.fg_override > 0 ? a.fg_override : (int)g.real(NumberOfFgObjects);
    if (fg > kPhiloxMaxObj) {  // more objects than the fixed strides hold: rendered without the rest, and the host is told
      if (a.truncated && k == kPhiloxMaxObj && lane == 0) *a.truncated = 1;
      fg = kPhiloxMaxObj;
    }
    int field_draws = 0;
    if (k == kPhiloxMaxObj) {
      // generateBackground, DG.cpp:2105-2143
      ofdg_blueprint bg = blank_bp();
      bg.obj_id = 1;
      bg.obj_type = OFDG_OBJ_POLYGON;
      bg.rot = g.trigger(BgRotTrigger) ? g.real(BgRot) : 0.f;
      bg.scale = g.trigger(BgScaleTrigger) ? g.real(BgScale) : 1.f;
      const float ptx = g.real(BgTransX), pty = g.real(BgTransY);
      bg.trans_x = cosf(-bg.rot) * ptx - sinf(-bg.rot) * pty;
      bg.trans_y = sinf(-bg.rot) * ptx + cosf(-bg.rot) * pty;
      bg.tex_id = g.integer(BgTexID);
      bg.tex_rot = g.real(BgInitRot);
      bg.tex_scale = g.real(BgInitScale);
      bg.tex_shift_x = g.integer(BgInitTransX);
      bg.tex_shift_y = g.integer(BgInitTransY);
      bg.do_warpfield_deformation = g.trigger(ObjDeformsNonrigidly);
      if (a.mode == 9 && bg.do_warpfield_deformation && a.n_fields > 0) {
        uint32_t r0, r1;
        g.rng.raw(FieldPick, r0, r1);
        bg.field_id = (int)(r0 % (uint32_t)a.n_fields);
      }
      if (lane == 0) {
        a.bp[s * kPhiloxMaxBp] = bg;
        a.n_top[s] = fg;
      }
      if (a.augment) {
        ofdg_augment au;
        au.enabled = 1;
        for (int c = 0; c < 3; ++c) au.gain[c] = 0.8f + 0.4f * g.rng.unit(AugGain);
        au.brightness = -20.f + 40.f * g.rng.unit(AugBrightness);
        au.contrast = 0.7f + 0.6f * g.rng.unit(AugContrast);
        au.noise_sigma = 10.f * g.rng.unit(AugSigma);
        g.rng.raw(AugSeed, au.noise_seed[0], au.noise_seed[1]);
        if (lane == 0) a.samples[s].aug = au;
      } else if (lane == 0) {
        a.samples[s].aug.enabled = 0;
      }
    } else if (k < fg) {
      // generateForegroundObject for object k: its own engines, its own slice of the arrays (staged in shared memory;
      // seg_begin / comp_begin / parent hold the absolute positions the slice will have in the batch's arrays)
      g.rng.init(a.seed, a.first_sample + (uint64_t)s, k);
      SampleOut o;
      o.bp_base = bp_base;
      o.seg_base = seg_base;
      o.bp = st.bp; o.seg_type = st.seg_type; o.seg_x = st.seg_x; o.seg_y = st.seg_y;
      o.nbp = 1; o.nseg = 0;
      {
        ofdg_blueprint nb = blank_bp();
        nb.obj_id = 10 + k;
        put_bp(o, 0, nb);
      }
      gen_object(g, o, 0, a.n_fields, field_draws);
      nbp = o.nbp; nseg = o.nseg;
    }
    if (k < kPhiloxMaxObj && lane == 0) { a.obj_nbp[s * kPhiloxMaxObj + k] = nbp; a.obj_nseg[s * kPhiloxMaxObj + k] = nseg; }
  }
  __syncwarp();
  {
    int* dst = reinterpret_cast<int*>(a.bp + bp_base);
    const int* src = reinterpret_cast<const int*>(st.bp);
    for (int i = lane; i < nbp * (int)(sizeof(ofdg_blueprint) / 4); i += 32) dst[i] = src[i];
    for (int i = lane; i < nseg; i += 32) {
      a.seg_type[seg_base + i] = st.seg_type[i];
      a.seg_x[seg_base + i] = st.seg_x[i];
      a.seg_y[seg_base + i] = st.seg_y[i];
    }
  }
}

// ---- device flatten (host/flatten.cpp restated for one thread per object) ------------------------------
struct VertOut {
  FlatVertex* v;
  int n, cap;
  int x0, y0, x1, y1;
  __device__ void begin(FlatVertex* p, int c) { v = p; n = 0; cap = c; x0 = y0 = 0x7FFFFFFF; x1 = y1 = -0x7FFFFFFF; }
  __device__ void push(double x, double y) {
    const int fx = iround(x * 256.0), fy = iround(y * 256.0);
    if (n < cap) { v[n].x = fx; v[n].y = fy; }
    ++n;
    x0 = min(x0, fx); x1 = max(x1, fx); y0 = min(y0, fy); y1 = max(y1, fy);
  }
};

// agg::curve3_div::recursive_bezier without recursion: pending right halves wait on an explicit stack
__device__ void subdivide(double x1, double y1, double x2, double y2, double x3, double y3, VertOut& out) {
  struct Seg { double x1, y1, x2, y2, x3, y3; int level; };
  Seg stack[34];
  int sp = 0;
  stack[sp++] = Seg{x1, y1, x2, y2, x3, y3, 0};
  while (sp > 0) {
    const Seg c = stack[--sp];
    if (c.level > 32) continue;
    const double x12 = (c.x1 + c.x2) / 2, y12 = (c.y1 + c.y2) / 2, x23 = (c.x2 + c.x3) / 2, y23 = (c.y2 + c.y3) / 2;
    const double x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2;
    const double dx = c.x3 - c.x1, dy = c.y3 - c.y1;
    double d = fabs(((c.x2 - c.x3) * dy - (c.y2 - c.y3) * dx));
    if (d > 1e-30) {
      if (d * d <= 0.25 * (dx * dx + dy * dy)) { out.push(x123, y123); continue; }
    } else {
      const double da = dx * dx + dy * dy;
      if (da == 0) d = (c.x2 - c.x1) * (c.x2 - c.x1) + (c.y2 - c.y1) * (c.y2 - c.y1);
      else {
        d = ((c.x2 - c.x1) * dx + (c.y2 - c.y1) * dy) / da;
        if (d > 0 && d < 1) continue;
        if (d <= 0) d = (c.x2 - c.x1) * (c.x2 - c.x1) + (c.y2 - c.y1) * (c.y2 - c.y1);
        else if (d >= 1) d = (c.x3 - c.x2) * (c.x3 - c.x2) + (c.y3 - c.y2) * (c.y3 - c.y2);
        else { const double ex = c.x1 + d * dx - c.x2, ey = c.y1 + d * dy - c.y2; d = ex * ex + ey * ey; }
      }
      if (d < 0.25) { out.push(c.x2, c.y2); continue; }
    }
    // left half first: push the right half, then the left one on top of it
    stack[sp++] = Seg{x123, y123, x23, y23, c.x3, c.y3, c.level + 1};
    stack[sp++] = Seg{c.x1, c.y1, x12, y12, x123, y123, c.level + 1};
  }
}

__device__ void outline(const PhiloxArgs& a, const ofdg_blueprint& b, const Affine& m, VertOut& out) {
  if (b.obj_type == OFDG_OBJ_ELLIPSE) {
    for (int st = 0; st < 100; ++st) {
      double x = 0.0 + c_circle_cos[st] * (double)b.ellipse_scale_x, y = 0.0 + c_circle_sin[st] * (double)b.ellipse_scale_y;
      m.apply(&x, &y);
      out.push(x, y);
    }
    return;
  }
  const int32_t* st = a.seg_type + b.seg_begin;
  const float* sx = a.seg_x + b.seg_begin;
  const float* sy = a.seg_y + b.seg_begin;
  double lx = sx[0], ly = sy[0];
  m.apply(&lx, &ly);
  out.push(lx, ly);
  for (int i = 1; i < b.seg_count; ++i) {
    if (st[i] == OFDG_SEG_CURVE3 && i + 1 < b.seg_count) {
      double cx = sx[i], cy = sy[i], ex = sx[i + 1], ey = sy[i + 1];
      m.apply(&cx, &cy);
      m.apply(&ex, &ey);
      subdivide(lx, ly, cx, cy, ex, ey, out);
      out.push(ex, ey);
      lx = ex; ly = ey;
      ++i;
    } else {
      lx = sx[i]; ly = sy[i];
      m.apply(&lx, &ly);
      out.push(lx, ly);
    }
  }
}

__device__ Affine motion_of(const ofdg_blueprint& b) {
  Affine m;
  m.then(Affine::rotation(b.rot));
  m.then(Affine::scaling(b.scale));
  m.then(Affine::translation(b.trans_x, b.trans_y));
  return m;
}

__device__ float cimg_mod_dev(float x, float m) {
  const double dx = (double)x, dm = (double)m;
  return (float)(dx - dm * floor(dx / dm));
}

__device__ void prepare_bg(const PhiloxArgs& a, const ofdg_blueprint& b, const Affine& tex_inv, int spread, BgPrep& p) {
  const int W = a.W, H = a.H, tw = 2 * W, th = 2 * H;
  p.tex = (int)((unsigned)b.tex_id % (unsigned)a.n_tex);
  const int w = a.tex_info[p.tex].w, h = a.tex_info[p.tex].h;
  p.shift_x = b.tex_shift_x; p.shift_y = b.tex_shift_y;
  const float nangle = cimg_mod_dev(b.tex_rot, 360.0f);
  p.rot_identity = (cimg_mod_dev(nangle, 90.0f) == 0) ? 1 : 0;
  if (p.rot_identity) { p.ca = 1.f; p.sa = 0.f; p.rw = w; p.rh = h; }
  else {
    const float rad = (float)(nangle * 3.14159265358979323846 / 180.0);
    p.ca = cosf(rad); p.sa = sinf(rad);
    const float ux = fabsf((unsigned)(w - 1) * p.ca), uy = fabsf((unsigned)(w - 1) * p.sa), vx = fabsf((unsigned)(h - 1) * p.sa), vy = fabsf((unsigned)(h - 1) * p.ca);
    p.rw = (int)floorf((1 + ux + vx) + 0.5f);
    p.rh = (int)floorf((1 + uy + vy) + 0.5f);
  }
  p.w2 = 0.5f * (unsigned)(w - 1); p.h2 = 0.5f * (unsigned)(h - 1);
  p.rw2 = 0.5f * (unsigned)(p.rw - 1); p.rh2 = 0.5f * (unsigned)(p.rh - 1);
  if (w >= tw && h >= th) {
    const float zoom = b.tex_scale;
    const int x0 = w / 2 - tw / 2, y0 = h / 2 - th / 2;
    const int x1 = (int)(w / 2 - tw / 2 + tw / zoom - 1), y1 = (int)(h / 2 - th / 2 + th / zoom - 1);
    p.crop_x0 = min(x0, x1); p.crop_y0 = min(y0, y1);
    p.crop_w = abs(x1 - x0) + 1; p.crop_h = abs(y1 - y0) + 1;
  } else {  // smaller than 2W x 2H: the whole rotated image is resized (DG.cpp:103-107)
    p.crop_x0 = 0; p.crop_y0 = 0; p.crop_w = p.rw; p.crop_h = p.rh;
  }
  p.crop_w = max(2, min(p.crop_w, 40 * tw)); p.crop_h = max(2, min(p.crop_h, 40 * th));  // the host path rejects these; stay in bounds here
  p.general = (p.crop_w * 10 > tw * 13 || p.crop_h * 10 > th * 13) ? 1 : 0;
  p.pad = 0;
  int nx0 = W / 2, ny0 = H / 2, nx1 = W / 2 + W - 1, ny1 = H / 2 + H - 1;
  if (spread < 0) { nx0 = 0; ny0 = 0; nx1 = tw - 1; ny1 = th - 1; }
  else {  // (host/flatten.cpp: a background with a warp field samples the warped canvas up to `spread` pixels outside the window)
    double fx0 = 1e300, fy0 = 1e300, fx1 = -1e300, fy1 = -1e300;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        double x = i ? fmin((double)tw, W / 2 + W + 1.0 + spread) : fmax(0.0, W / 2 + 0.0 - spread);
        double y = j ? fmin((double)th, H / 2 + H + 1.0 + spread) : fmax(0.0, H / 2 + 0.0 - spread);
        tex_inv.apply(&x, &y);
        fx0 = fmin(fx0, x); fx1 = fmax(fx1, x); fy0 = fmin(fy0, y); fy1 = fmax(fy1, y);
      }
    const int ax0 = (int)floor(fx0) - 3, ax1 = (int)ceil(fx1) + 3, ay0 = (int)floor(fy0) - 3, ay1 = (int)ceil(fy1) + 3;
    if (ax0 < 0 || ax1 > tw - 1) { nx0 = 0; nx1 = tw - 1; } else { nx0 = min(nx0, ax0); nx1 = max(nx1, ax1); }
    if (ay0 < 0 || ay1 > th - 1) { ny0 = 0; ny1 = th - 1; } else { ny0 = min(ny0, ay0); ny1 = max(ny1, ay1); }
  }
  p.need[0] = nx0; p.need[1] = ny0; p.need[2] = nx1; p.need[3] = ny1;
}

constexpr int kFlatObjPerBlock = 8, kFlatLanes = 2 * kPhiloxMaxShapes;  // 16 (outline, frame) threads per object

__global__ void __launch_bounds__(kFlatObjPerBlock * kFlatLanes) philox_flatten_kernel(PhiloxArgs a) {
  __shared__ int s_box[kFlatObjPerBlock][2][4];
  const int s = blockIdx.x;
  const int W = a.W, H = a.H;
  const ofdg_blueprint* bp = a.bp;
  const ofdg_blueprint& bg = bp[s * kPhiloxMaxBp];
  const Affine bgM = motion_of(bg);
  if (blockIdx.y == kPhiloxMaxObj / kFlatObjPerBlock) {  // the extra block row: background / FlatSample
    if (threadIdx.x != 0) return;
    FlatSample smp = a.samples[s];  // keeps the augmentation record written by the parameter kernel
    Affine bgI;
    bgI.then(Affine::rotation(0.0));
    bgI.then(Affine::translation((double)W, (double)H));
    Affine tex_tf = bgI.inverse();
    tex_tf.then(bgM);
    tex_tf.then(bgI);
    const Affine tex_inv = tex_tf.inverse();
    tex_inv.store(smp.bg_tex_inv);
    bgM.store(smp.bg_motion);
    bgM.inverse().store(smp.bg_motion_inv);
    smp.bg_field = (a.mode == 9 && bg.do_warpfield_deformation && bg.field_id >= 0) ? bg.field_id : -1;
    const int spread = smp.bg_field < 0 ? 0 : (a.field_reach ? 2 * min(a.field_reach[smp.bg_field], 4 * (W + H)) + 2 : -1);
    prepare_bg(a, bg, tex_inv, spread, smp.prep);
    smp.obj_begin = s * kPhiloxMaxObj;
    smp.obj_count = a.n_top[s];
    a.samples[s] = smp;
    return;
  }
  const int lo = threadIdx.x / kFlatLanes, sf = threadIdx.x % kFlatLanes, si = sf >> 1, f = sf & 1;
  const int k = blockIdx.y * kFlatObjPerBlock + lo;
  if (sf < 8) s_box[lo][sf >> 2][sf & 3] = (sf & 2) ? -0x7FFFFFFF : 0x7FFFFFFF;
  __syncthreads();
  const bool active = k < a.n_top[s];
  const int obj_slot = s * kPhiloxMaxObj + k;
  const ofdg_blueprint& b = bp[s * kPhiloxMaxBp + 1 + k * kPhiloxMaxShapes];
  const bool composite = active && b.obj_type == OFDG_OBJ_COMPOSITE;
  const int nshape = !active ? 0 : (composite ? b.comp_count : 1);
  Affine bg_n = Affine::translation(-W / 2., -H / 2.);
  bg_n.then(bgM);
  bg_n.then(Affine::translation(W / 2., H / 2.));
  if (si < nshape) {
    const ofdg_blueprint& c = composite ? bp[b.comp_begin + si] : b;
    Affine T;
    T.then(Affine::rotation(c.init_rot));
    T.then(Affine::translation(c.init_trans_x, c.init_trans_y));
    if (f) {
      Affine Mc = motion_of(c);
      Mc.then(bg_n);
      T.then(Mc);
    }
    VertOut out;
    const int cap = kPhiloxMaxVerts / kFlatLanes;
    const size_t vb = (size_t)obj_slot * kPhiloxMaxVerts + (size_t)sf * cap;
    out.begin(a.verts + vb, cap);
    outline(a, c, T, out);
    if (out.n > out.cap) {  // out of room: drop the outline rather than corrupt memory, and tell the host (never seen with the modes' shapes)
      out.n = 0;
      if (a.truncated) *a.truncated = 1;
    }
    FlatShape& sh = a.shapes[obj_slot * kPhiloxMaxShapes + si];
    sh.vbegin[f] = (int)vb;
    sh.vcount[f] = out.n;
    int bx[4] = {0x7FFFFFF0, 0x7FFFFFF0, -0x7FFFFFF0, -0x7FFFFFF0};
    if (out.n) { bx[0] = out.x0 >> 8; bx[1] = out.y0 >> 8; bx[2] = out.x1 >> 8; bx[3] = out.y1 >> 8; }
    for (int i = 0; i < 4; ++i) sh.bbox[f][i] = bx[i];
    if (f) {
      for (int i = 0; i < 4; ++i) sh.raw1[i] = bx[i];
      // MovingObjectBase::renderMasks warps this outline's frame-1 masks by the inverse field (DG.cpp:370-386); components
      // carry their parent's field (DG.cpp:1157-1163). The slot order is arbitrary: slots only name scratch planes.
      const int field = (a.mode == 9 && b.do_warpfield_deformation && b.field_id >= 0) ? b.field_id : -1;
      int slot = -1;
      if (field >= 0) {
        slot = atomicAdd(a.n_deform, 1);
        a.deform_shape[slot] = obj_slot * kPhiloxMaxShapes + si;
        a.deform_field[slot] = field;
        const int reach = a.field_reach[field] + 2;
        bx[0] -= reach; bx[1] -= reach; bx[2] += reach; bx[3] += reach;
        for (int i = 0; i < 4; ++i) sh.bbox[1][i] = bx[i];
      }
      sh.deform = slot;
    } else {
      sh.additive = c.is_additive_component ? 1 : 0;
    }
    atomicMin(&s_box[lo][f][0], bx[0]); atomicMin(&s_box[lo][f][1], bx[1]);
    atomicMax(&s_box[lo][f][2], bx[2]); atomicMax(&s_box[lo][f][3], bx[3]);
  }
  __syncthreads();
  if (active && sf == 0) {
    FlatObject o;
    memset(&o, 0, sizeof(o));
    o.obj_id = b.obj_id;
    o.tex = (int)((unsigned)b.tex_id % (unsigned)a.n_tex);
    o.field = (a.mode == 9 && b.do_warpfield_deformation && b.field_id >= 0) ? b.field_id : -1;
    o.composite = composite;
    o.shape_begin = obj_slot * kPhiloxMaxShapes;
    o.shape_count = nshape;
    for (int ff = 0; ff < 2; ++ff)
      for (int i = 0; i < 4; ++i) o.bbox[ff][i] = s_box[lo][ff][i];
    Affine M = motion_of(b);
    M.then(bg_n);
    M.store(o.motion);
    M.inverse().store(o.tex_inv);
    a.objects[obj_slot] = o;
  }
}

}  // namespace

void philox_upload_circle(const double* c, const double* s) {
  cudaMemcpyToSymbol(c_circle_cos, c, 100 * sizeof(double));
  cudaMemcpyToSymbol(c_circle_sin, s, 100 * sizeof(double));
}

int launch_philox(const PhiloxArgs& a, cudaStream_t s) {
  philox_params_kernel<<<dim3(a.batch, kParamBlocks), 32 * kParamWarps, 0, s>>>(a);
  philox_flatten_kernel<<<dim3(a.batch, kPhiloxMaxObj / kFlatObjPerBlock + 1), kFlatObjPerBlock * kFlatLanes, 0, s>>>(a);
  return 2;
}

}  // namespace ofdg
