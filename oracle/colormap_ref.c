/* oracle/colormap_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Restates the colour ramp of SpecCache::populateTex, reference spec-cache.cpp:77-96.
 * Note the integer constants: 255/3 == 85 and 2*255/3 == 170 (int arithmetic in the reference).
 * PINNED against the reference itself: spec-cache.cpp compiled unmodified (oracle/_ref/libapp_ref.so,
 * the GL shim records the texels handed to glTexImage1D) uploads the same bytes (tests/test_oracle.py). */
#include "oracle.h"
#include <math.h>

void mlxo_colormap(const float *spec, int count, float k, uint8_t *rgb) {
  for (int i = 0; i < count; ++i) {
    float tmp = spec[i] * k; /* spec-cache.cpp:79  std::clamp(s[i]*k, 0.f, 255.f) */
    if (tmp < 0.f) tmp = 0.f;
    if (tmp > 255.f) tmp = 255.f;
    uint8_t r, g, b;
    if (tmp < 255 / 3) { /* :80-83 */
      r = (uint8_t)tmp;
      g = 0;
      b = 0;
    } else if (tmp < 2 * 255 / 3) { /* :84-90  (tmp-85) float, /85 float, *3.141592 double */
      const double a = (tmp - 255 / 3) / (255 / 3) * 3.141592 / 2;
      r = (uint8_t)(tmp * cos(a));
      g = (uint8_t)(tmp * sin(a));
      b = 0;
    } else { /* :91-95 */
      const uint8_t lk = (uint8_t)((tmp - 2 * 255 / 3) * 3);
      r = lk;
      g = (uint8_t)tmp;
      b = lk;
    }
    rgb[3 * i] = r;
    rgb[3 * i + 1] = g;
    rgb[3 * i + 2] = b;
  }
}
