/*
 * ofdg/augment.h -- the arithmetic of the colour/noise augmentation (specification in ofdg/scene.h),
 * shared verbatim by the sm_100a render kernel and by the CPU oracle: integer Philox4x32-10 and
 * single-rounded float operations only, so both sides agree bit for bit. Not part of the reference.
 */
#ifndef OFDG_AUGMENT_H_
#define OFDG_AUGMENT_H_
#include <stdint.h>

#include "ofdg/scene.h"

#if defined(__CUDACC__)
#define OFDG_AUG_FN __host__ __device__ __forceinline__
#else
#define OFDG_AUG_FN static inline
#endif

/* Noise sums of pixel p in frame f for the three channels, from ONE Philox4x32-10(key = noise_seed, counter = {p, f, 0, 0}):
 * channel c gets the sum of the four bytes of output word c (0..1020: Irwin-Hall of four bytes, mean 510, sigma 147.80),
 * packed ten bits per channel. (Revision 2 of the specification: revision 1 ran one Philox per channel and summed eight 16-bit
 * halves -- three times the integer work per pixel for a noise term nobody can tell apart.) */
OFDG_AUG_FN uint32_t ofdg_noise3(uint32_t k0, uint32_t k1, uint32_t p, uint32_t f) {
  uint32_t x0 = p, x1 = f, x2 = 0u, x3 = 0u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
    const uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ k0, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ x3 ^ k1, y3 = (uint32_t)p0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
#if defined(__CUDA_ARCH__)
  const uint32_t s0 = __dp4a(x0, 0x01010101u, 0u), s1 = __dp4a(x1, 0x01010101u, 0u), s2 = __dp4a(x2, 0x01010101u, 0u);
#else
  const uint32_t s0 = (x0 & 255u) + ((x0 >> 8) & 255u) + ((x0 >> 16) & 255u) + (x0 >> 24);
  const uint32_t s1 = (x1 & 255u) + ((x1 >> 8) & 255u) + ((x1 >> 16) & 255u) + (x1 >> 24);
  const uint32_t s2 = (x2 & 255u) + ((x2 >> 8) & 255u) + ((x2 >> 16) & 255u) + (x2 >> 24);
#endif
  return s0 | (s1 << 10) | (s2 << 20);
}

/* One output value from the composited 8-bit value v of channel c and the pixel's packed noise sums */
OFDG_AUG_FN float ofdg_augment_apply(const ofdg_augment* a, float v, int c, uint32_t noise3) {
  const float n = ((float)(int32_t)((noise3 >> (10 * c)) & 1023u) - 510.0f) * (1.0f / 147.80f);
  float y = a->gain[c] * v;
  y = y - 127.5f;
  y = a->contrast * y;
  y = y + 127.5f;
  y = y + a->brightness;
  y = y + a->noise_sigma * n;
  y = y < 0.f ? 0.f : y;
  return y > 255.f ? 255.f : y;
}

/* The same, one value at a time: c = channel 0..2, f = frame 0/1, p = pixel index y*W+x */
OFDG_AUG_FN float ofdg_augment_value(const ofdg_augment* a, float v, int c, int f, uint32_t p) {
  return ofdg_augment_apply(a, v, c, ofdg_noise3(a->noise_seed[0], a->noise_seed[1], p, (uint32_t)f));
}

#endif
