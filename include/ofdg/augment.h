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

/* Sum of the eight 16-bit halves of Philox4x32-10(key, counter) - 262140: Irwin-Hall, mean 0, sigma 53510.1 */
OFDG_AUG_FN int32_t ofdg_noise_sum(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1) {
  uint32_t x0 = c0, x1 = c1, x2 = 0u, x3 = 0u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
    const uint32_t y0 = (uint32_t)(p1 >> 32) ^ x1 ^ k0, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ x3 ^ k1, y3 = (uint32_t)p0;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint32_t s = (x0 & 0xFFFFu) + (x0 >> 16) + (x1 & 0xFFFFu) + (x1 >> 16) + (x2 & 0xFFFFu) + (x2 >> 16) + (x3 & 0xFFFFu) + (x3 >> 16);
  return (int32_t)s - 262140;
}

/* One output value: v = composited 8-bit value, c = channel 0..2, f = frame 0/1, p = pixel index y*W+x */
OFDG_AUG_FN float ofdg_augment_value(const ofdg_augment* a, float v, int c, int f, uint32_t p) {
  const float n = (float)ofdg_noise_sum(a->noise_seed[0], a->noise_seed[1], p, (uint32_t)(2 * c + f)) * (1.0f / 53510.1f);
  float y = a->gain[c] * v;
  y = y - 127.5f;
  y = a->contrast * y;
  y = y + 127.5f;
  y = y + a->brightness;
  y = y + a->noise_sigma * n;
  y = y < 0.f ? 0.f : y;
  return y > 255.f ? 255.f : y;
}

#endif
