// melonix_b200/csrc/spec_frame.cuh -- per-thread pieces of the regular-hop Spec kernel (K1r).
//
// One frame of Spec::internalGetSpec (reference spec.cpp:44-66) is transformed by a group of
// TPF = N/32 threads as an N/2-point complex FFT of the even/odd packed real frame.  The three
// per-thread steps around the FFT live here as host/device functions so that the same code is
// compiled by g++ and run under a sequential emulation of the thread group against the oracle
// (tests/host/spec_frame_emul.cpp) -- the index arithmetic is verified without a GPU.
#pragma once
#include <math.h>

#include "fft.cuh"

namespace mlx {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float spec_sqrt(float x) {  // sqrt.approx: ~1 ulp, far inside the 1e-4 RMS budget
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float spec_mul(float a, float b) { return __fmul_rn(a, b); }
// a product the compiler must not hoist out of the frame loop: the derived window factors are loop-invariant, and
// hoisted they are 32 registers per thread again (spilled: the local-memory loads cost what the table loads did)
__device__ __forceinline__ float spec_mul_here(float a, float b) {
  float r;
  asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
#else
inline float spec_sqrt(float x) { return __builtin_sqrtf(x); }
inline float spec_mul(float a, float b) { return a * b; }
inline float spec_mul_here(float a, float b) { return a * b; }
#endif

#ifndef MLX_SPEC_DERIVE
#define MLX_SPEC_DERIVE 1  // K1r: window factors and pair-split twiddles of a thread's slots derived from ONE loaded value
                           // by compile-time constant factors instead of one shared-memory load each (the kernel runs
                           // at 86 % of the L1 / shared-memory data pipe and 64 % of the issue slots: loads are dearer)
#endif

// exp() and cos/sin(2 pi q / 32) as compile-time constants (Taylor series in double; arguments are small)
MLX_HDC double spec_cexp(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 40; ++i) {
    term *= x / i;
    sum += term;
  }
  return sum;
}
MLX_HDC double spec_csin(double x) {
  double term = x, sum = x;
  for (int i = 1; i < 20; ++i) {
    term *= -x * x / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
MLX_HDC double spec_ccos(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 20; ++i) {
    term *= -x * x / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}

struct SpecRot32 {  // exp(-2 pi i q / 32), q < 8
  float c[8], s[8];
};
MLX_HDC SpecRot32 spec_rot32() {
  SpecRot32 r{};
  for (int q = 0; q < 8; ++q) {
    r.c[q] = (float)spec_ccos(6.283185307179586476925286766559 * q / 32.0);
    r.s[q] = (float)-spec_csin(6.283185307179586476925286766559 * q / 32.0);
  }
  return r;
}
struct SpecGeo16 {  // exp(2.5e-4f * step * m), m < 16
  float k[16];
};
MLX_HDC SpecGeo16 spec_geo16(int step) {
  SpecGeo16 g{};
  for (int m = 0; m < 16; ++m) g.k[m] = (float)spec_cexp((double)2.5e-4f * step * m);  // the reference's float constant
  return g;
}

template <int N>
struct SpecFrame {
  static constexpr int NC = N / 2;
  static constexpr int TPF = NC / 16;
  using C = cplx<float>;
  using F = Fft<float, NC, -1>;

  // Window [end-N, end) of the frame, already in `cur` (zero outside the track): slot m takes the
  // samples p = 2(t + m*TPF), p + 1, each multiplied by its window factor -- expf(-2.5e-4f*(start-i))
  // before `start`, 1 from `start` on (spec.cpp:55-58; the float product is the reference's own).
  // win2(p) returns the factors of positions p and p + 1.
  template <class Win2>
  static MLX_HD void load(C (&x)[16], const float* cur, int t, Win2&& win2) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int p = 2 * (t + m * TPF);
      const C s2 = *reinterpret_cast<const C*>(cur + p);
      const C w2 = win2(p);
      x[m] = C{spec_mul(w2.x, s2.x), spec_mul(w2.y, s2.y)};
    }
  }

  // The same load with the window factors DERIVED: before `start` the window is the geometric sequence
  // expf(-2.5e-4f * (start - i)), and slot m of a thread sits m * 2 TPF samples after slot 0, so
  //   w(slot m) = min(1, w(slot 0) * exp(2.5e-4f * 2 TPF m))
  // -- the min() supplies the factor 1 from `start` on, where the continued sequence exceeds 1.  c0 / c1 are the
  // (exact, host-table) factors of the thread's two samples of slot 0, or 1 when those lie at or after `start`;
  // the derived factors differ from the table's by a few ulp (the results by ~1e-8 of their scale).
  static MLX_HD void load_derived(C (&x)[16], const float* cur, int t, float c0, float c1) {
    constexpr SpecGeo16 geo = spec_geo16(2 * TPF);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float km = geo.k[m];
      const int p = 2 * (t + m * TPF);
      const C s2 = *reinterpret_cast<const C*>(cur + p);
      x[m] = C{spec_mul(fminf(1.f, spec_mul_here(c0, km)), s2.x), spec_mul(fminf(1.f, spec_mul_here(c1, km)), s2.y)};
    }
  }

  // After the FFT x[m] = Z[t + m*TPF].  Bin k = t + q*TPF (q < 8) pairs with bin NC - k, which is
  // slot 15 - q (16 - q for t = 0) of thread (TPF - t) mod TPF: only the upper slots are exchanged.
  // They are staged UNPADDED, slot m at (m - 8)*TPF + t: Z[NC - k] is then element (8 - q)*TPF - t, so the
  // lanes of a warp read consecutive 8-byte words in descending order -- one wavefront per half-warp (in the
  // padded transform layout the descending run crosses a padding element: two-way conflicts, 16 of the 284
  // shared-memory wavefronts of a 1024-point frame).  The staging area overlaps other threads' transform
  // slots: the caller synchronises the group between the last stage's loads and stage_upper().
  static MLX_HD void stage_upper(const C (&x)[16], C* buf, int t) {
    C* p = buf + t;
#pragma unroll
    for (int m = 8; m < 16; ++m) p[(m - 8) * TPF] = x[m];
  }

  // |X[k]| / N for the bins of this thread: X[k] = E[k] + W^k O[k] from Z[k] and Z[NC-k]; the split
  // works on 2X, `scale` = 0.5 / N.  emit(k, value) is called for every k in [0, NC) exactly once
  // over the group; the Nyquist bin X[NC] is dropped as the reference does (spec.cpp:61).
  // DERIVE: twr(t) is loaded once by the caller (`w0`), the twiddle of bin t + q*TPF is w0 * exp(-2 pi i q / 32)
  // (TPF / N = 1 / 32 for every N) with the rotation as compile-time constants.
  template <bool DERIVE = false, class Twr, class Emit>
  static MLX_HD void emit_bins(const C (&x)[16], const C* buf, int t, float scale, Twr&& twr, Emit&& emit,
                               C w0 = C{1.f, 0.f}) {
    constexpr SpecRot32 rot = spec_rot32();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int k = t + q * TPF;
      const C za = x[q];
      C zc = buf[(8 - q) * TPF - t];  // q == 0, t == 0 reads one element past the staging area (inside the buffer)
      if (q == 0 && t == 0) zc = za;  // Z[NC] == Z[0] (slot 0 is not staged)
      C w;                            // exp(-2 pi i k / N)
      if constexpr (DERIVE) {
        const float rc = rot.c[q], rs = rot.s[q];
        w = q == 0 ? w0 : C{w0.x * rc - w0.y * rs, w0.x * rs + w0.y * rc};
      } else {
        w = twr(k);
      }
      const float er = za.x + zc.x, ei = za.y - zc.y;
      const float dr = za.x - zc.x, di = za.y + zc.y;
      const float tr_ = dr * w.x - di * w.y, ti_ = dr * w.y + di * w.x;
      const float xkr = er + ti_, xki = ei - tr_;
      const float xmr = er - ti_, xmi = -ei - tr_;
      emit(k, spec_sqrt(fmaf(xkr, xkr, xki * xki)) * scale);
      if (q != 0 || t != 0) emit(NC - k, spec_sqrt(fmaf(xmr, xmr, xmi * xmi)) * scale);
    }
    // |X[NC/2]| = |Z[NC/2]|, slot 8 of thread 0
    if (t == 0) emit(NC / 2, spec_sqrt(fmaf(x[8].x, x[8].x, x[8].y * x[8].y)) * (2.f * scale));
  }
};

}  // namespace mlx
