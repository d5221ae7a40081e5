"""Revision 2 of this repository's augmentation specification (include/ofdg/augment.h, include/ofdg/scene.h; not part of the
reference): the header is compiled for the host with gcc and pinned by known answers and by the statistics of its noise term.
The same header is what the sm_100a blob write and the CPU oracle evaluate (tests/test_gpu_parity.py::test_augmentation_bit_exact)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <stdio.h>
#include "ofdg/augment.h"
int main(void) {
  printf("%u %u %u\n", ofdg_noise3(0u, 0u, 0u, 0u), ofdg_noise3(0x12345678u, 0x9abcdef0u, 12345u, 1u), ofdg_noise3(1u, 2u, 196607u, 0u));
  double s[3] = {0, 0, 0}, q[3] = {0, 0, 0}, c01 = 0;
  const int n = 200000;
  for (int p = 0; p < n; ++p) {
    const uint32_t w = ofdg_noise3(777u, 888u, (uint32_t)p, (uint32_t)(p & 1));
    double v[3];
    for (int c = 0; c < 3; ++c) { v[c] = ((double)((w >> (10 * c)) & 1023u) - 510.0) / 147.80; s[c] += v[c]; q[c] += v[c] * v[c]; }
    c01 += v[0] * v[1];
  }
  for (int c = 0; c < 3; ++c) printf("%.6f %.6f\n", s[c] / n, q[c] / n);
  printf("%.6f\n", c01 / n);
  ofdg_augment a = {1, {1.0f, 1.1f, 0.9f}, 3.0f, 1.2f, 5.0f, {42u, 43u}};
  printf("%.9g %.9g %.9g\n", ofdg_augment_value(&a, 100.0f, 0, 0, 7u), ofdg_augment_value(&a, 200.0f, 1, 1, 8u), ofdg_augment_value(&a, 0.0f, 2, 0, 9u));
  return 0;
}
"""


def test_augmentation_spec_known_answers_and_noise_statistics():
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "a.c"), os.path.join(d, "a")
        open(c, "w").write(SRC)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-o", exe, c], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    # one Philox4x32-10 per (pixel, frame): three ten-bit byte sums packed into one word
    assert out[0].split() == ["549089866", "563613357", "645402181"]
    for c in range(3):  # Irwin-Hall of four bytes, scaled: mean 0, variance 1, channels uncorrelated
        mean, var = map(float, out[1 + c].split())
        assert abs(mean) < 0.01 and abs(var - 1.0) < 0.01
    assert abs(float(out[4])) < 0.01
    # y = contrast * (gain * v - 127.5) + 127.5 + brightness + sigma * n, clamped to [0, 255], one rounding per operation
    assert out[5].split() == ["90.5649567", "236.696213", "0"]
