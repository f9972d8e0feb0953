// melonix_b200/csrc/pv_analysis.cuh -- one analysis bin of the phase vocoder (PV-spec A.2-A.3): magnitude,
// integer-turn phase, wrapped phase advance and the FP64 decision at the +-pi cut.  Host/device so that
// tests/host/pv_analysis_emul.cpp can check the cut decisions against a double-precision evaluation of
// the spec without a GPU (the host build replaces the SFU approximations by IEEE division / sqrt).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "pv_shift.cuh"  // MagD, MLX_HD

namespace mlx {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pv_f2u_rn(float x) { return __float2uint_rn(x); }
__device__ __forceinline__ uint32_t pv_f2bits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ int pv_d2hi(double x) { return __double2hiint(x); }
#else  // host build (tests/host/pv_analysis_emul.cpp): IEEE division / sqrt instead of the SFU approximations
inline float fast_rcp(float x) { return 1.0f / x; }
inline float fast_sqrt(float x) { return sqrtf(x); }
inline uint32_t pv_f2u_rn(float x) { return (uint32_t)llrintf(x); }
inline uint32_t pv_f2bits(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return u;
}
inline int pv_d2hi(double x) {
  unsigned long long u;
  memcpy(&u, &x, 8);
  return (int)(u >> 32);
}
#endif

// |atan2(y, x)| in [0, pi] from |y| and signed x.  atan(t) = t * P(t^2) on [0,1] (degree-8 fit,
// 1.1e-7 rad max error evaluated in float), one MUFU.RCP, no branches, no slow paths.
MLX_HD float atan2_abs(float ay, float x) {
  const float ax = fabsf(x);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = mn * fast_rcp(mx);
  const float s = t * t;
  float p = 2.834064187e-03f;
  p = fmaf(p, s, -1.600502990e-02f);
  p = fmaf(p, s, 4.258760810e-02f);
  p = fmaf(p, s, -7.495445758e-02f);
  p = fmaf(p, s, 1.063675433e-01f);
  p = fmaf(p, s, -1.420257092e-01f);
  p = fmaf(p, s, 1.999248415e-01f);
  p = fmaf(p, s, -3.333306611e-01f);
  p = fmaf(p, s, 1.0f);
  p *= t;
  p = (ay > ax) ? 1.5707963267948966f - p : p;
  p = (x < 0.f) ? 3.1415926535897931f - p : p;
  return p;
}

// One analysis bin (PV-spec A.2-A.3).  X = (a, b) this frame and (c, d) previous frame in double.
//
// The frame's absolute phase is quantised to integer turns P = round(arg(X) / 2pi * 2^32) (float
// atan2 polynomial: its smooth error e(phi) enters the phase *difference* as e(phi_f) - e(phi_{f-1})
// and therefore telescopes over frames instead of accumulating), and
//     d = P_f - P_{f-1} - bin * 2^30   (mod 2^32, as a signed 32-bit number)
// is exact integer arithmetic.  The only discontinuous decision -- on which side of the +-pi cut d
// lies -- is taken from the sign of Im(X conj(Xprev) (-i)^bin) evaluated in DOUBLE; when the
// integer difference landed on the other side, `flip` tells the consumer to add -+2^32.
// step 1 of analysis_bin: the frame's own magnitude and integer-turn phase (no dependence on the previous frame)
MLX_HD void analysis_polar(double a, double b, float& mag, uint32_t& P) {
  const float af = (float)a, bf = (float)b;
  mag = fast_sqrt(fmaf(af, af, bf * bf));
  const float pabs = atan2_abs(fabsf(bf), af);
  P = pv_f2u_rn(pabs * 683565275.5764316f);  // 2^32 / (2 pi); pabs <= pi -> <= 2^31
  P = (pv_f2bits(bf) >> 31) ? (0u - P) : P;
}

// step 2: the wrapped phase advance against the previous frame (X_prev = (c, d), its phase and magnitude)
MLX_HD void analysis_advance(double a, double b, double c, double d, uint32_t P, uint32_t p_prev, float mag,
                             float mag_prev, int bin, bool real_bin, int& d32, bool& flip) {
  d32 = (int)(P - p_prev - ((uint32_t)bin << 30));
  const bool gate = mag * mag_prev <= 1e-18f;  // silence gate |Z| <= 1e-18 -> d = 0
  const unsigned dneg = (unsigned)d32 >> 31;
  const unsigned dabs = dneg ? (0u - (unsigned)d32) : (unsigned)d32;
  flip = false;
  // Only within 2^20 counts (1.5e-3 rad) of the +-pi cut can the float phases put d on the wrong
  // side (their error is ~1e-7 rad); only there the DOUBLE product decides.  Component of
  // X conj(Xprev) (-i)^bin that ends up as Im Z: odd bins -> -(ac + bd), even bins -> (bc - ad).
  if (!gate && dabs > 0x7FF00000u) {
    const bool odd = bin & 1, neg = bin & 2;
    const double u = odd ? a : b, v = odd ? b : -a;
    const double s64 = fma(u, c, v * d);
    unsigned sbit = ((unsigned)pv_d2hi(s64) >> 31) ^ (odd ? 1u : 0u) ^ (neg ? 1u : 0u);
    if (real_bin) sbit = 0u;  // bins 0 and N/2 are purely real: Im Z := +0, d in {0, +pi}
    flip = dneg != sbit;
  }
  d32 = gate ? 0 : d32;
}

MLX_HD void analysis_bin(double a, double b, double c, double d, uint32_t& p_prev,
                                             float& mag_prev, int bin, bool real_bin, float& mag, int& d32,
                                             bool& flip) {
  uint32_t P;
  analysis_polar(a, b, mag, P);
  analysis_advance(a, b, c, d, P, p_prev, mag, mag_prev, bin, real_bin, d32, flip);
  p_prev = P;
  mag_prev = mag;
}

// The purely real bins 0 and N/2: arg X is 0 or pi, Im Z := +0 (PV-spec v1), so d is 0 or +pi.
// Same result as analysis_bin(real_bin = true) at a fraction of its cost.
MLX_HD MagD analysis_real_bin(double x, uint32_t& p_prev, float& mag_prev) {
  const float mag = fabsf((float)x);
  const uint32_t P = x < 0.0 ? 0x80000000u : 0u;
  const bool gate = mag * mag_prev <= 1e-18f;
  const bool turned = (P != p_prev) && !gate;  // d = +pi: stored as -2^31 with the flip flag set
  p_prev = P;
  mag_prev = mag;
  return MagD{turned ? -mag : mag, turned ? (int)0x80000000u : 0};
}

}  // namespace mlx
