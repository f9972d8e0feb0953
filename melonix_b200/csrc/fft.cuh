// melonix_b200/csrc/fft.cuh
//
// Batched power-of-two complex FFT held entirely in shared memory + registers (no cuFFT).
// Stockham autosort, decimation in time.  A "group" of TPF = NC/16 threads transforms one frame of
// NC complex points; every thread keeps 16 points in registers, so each stage is one radix-16 (or,
// for the last stage, 16/R radix-R) butterfly per thread followed by an exchange through the
// frame's shared-memory buffer.  Stage plan: log2(NC) = 4a + b  ->  a radix-16 stages, then one
// radix-2^b stage.  The last stage is in place (reads and writes the same thread-private slots).
//
// Register slot m of thread t always corresponds to element index (t + m*TPF): that is the layout
// of the input handed to run() and of the natural-order output it returns.
//
// Shared-memory layout: array-of-complex (8 B for float, 16 B for double) with one padding element
// every 16 (pad(i) = i + i/16).  With 64/128-bit accesses issued per half/quarter warp this makes
// the stride-16 stores of the radix-16 stages and all unit-stride loads bank-conflict free.
//
// The file is also compiled by g++ (tests/host/fft_emul.cpp) with a sequential emulation of the
// thread group, which is how the index arithmetic is verified without a GPU.
#pragma once

#ifdef __CUDACC__
#define MLX_HD __host__ __device__ __forceinline__
#define MLX_D __device__ __forceinline__
#define MLX_HDC __host__ __device__ constexpr
#else
#define MLX_HD inline
#define MLX_D inline
#define MLX_HDC constexpr
#endif

#ifndef MLX_FFT_DERIVE_LAST
#define MLX_FFT_DERIVE_LAST 1  // double-precision transforms: the last stage keeps ONE twiddle per thread; butterfly b
                               // uses it rotated by the constant exp(DIR 2 pi i b / 16) (4 operations) instead of
                               // 16/R twiddle registers (K_A: 12 registers, 176 -> 88 bytes of spills, 14.05 -> 13.5 ms)
#endif
#ifndef MLX_FFT_PACKED_F32
#define MLX_FFT_PACKED_F32 1  // single-precision complex add / subtract as ONE packed FADD2 (sm_100a): K1r 1.106 -> 1.037 ms at
                              // 1024/256, K_S 5.71 -> 5.65 ms; FFMA2 runs at half the scalar issue rate (tools/ubench/fp32x2_probe.cu),
                              // so only the issue slots are saved, not pipe cycles
#endif
#ifndef MLX_FFT_TREE64
#define MLX_FFT_TREE64 0  // 1: double-precision twiddle powers by the product tree too (depth 4 instead of a chain of 14)
#endif

namespace mlx {

template <typename T>
struct alignas(2 * sizeof(T)) cplx {
  T x, y;
};

template <typename T>
MLX_HD cplx<T> cmul(const cplx<T> a, const cplx<T> b) {
  return cplx<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
MLX_HD cplx<T> cadd(const cplx<T> a, const cplx<T> b) {
  return cplx<T>{a.x + b.x, a.y + b.y};
}
template <typename T>
MLX_HD cplx<T> csub(const cplx<T> a, const cplx<T> b) {
  return cplx<T>{a.x - b.x, a.y - b.y};
}
#if defined(__CUDA_ARCH__) && MLX_FFT_PACKED_F32
// sm_100a packed single precision: a complex float is one aligned 64-bit register pair, so a complex
// add / subtract is ONE FADD2 (the negation folds into the operand modifier).  Same rounding as two FADDs.
// (MLX_FFT_PACKED_F32 = 2 / 3: only the additions / only the subtractions packed -- tuning experiments)
#if MLX_FFT_PACKED_F32 != 3
template <>
MLX_HD cplx<float> cadd<float>(const cplx<float> a, const cplx<float> b) {
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return cplx<float>{r.x, r.y};
}
#endif
#if MLX_FFT_PACKED_F32 != 2
template <>
MLX_HD cplx<float> csub<float>(const cplx<float> a, const cplx<float> b) {
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
  return cplx<float>{r.x, r.y};
}
#endif
#endif
// multiply by DIR*i  (DIR = -1: forward transform, e^{-i...};  DIR = +1: inverse)
template <int DIR, typename T>
MLX_HD cplx<T> cmul_i(const cplx<T> a) {
  return DIR < 0 ? cplx<T>{a.y, -a.x} : cplx<T>{-a.y, a.x};
}

MLX_HDC int fft_pad(int i) { return i + (i >> 4); }

// ---------------------------------------------------------------- small in-register DFTs
template <int DIR, typename T>
MLX_HD void dft2(cplx<T>& a, cplx<T>& b) {
  const cplx<T> s = cadd(a, b), d = csub(a, b);
  a = s;
  b = d;
}

template <int DIR, typename T>
MLX_HD void dft4(cplx<T>& a0, cplx<T>& a1, cplx<T>& a2, cplx<T>& a3) {
  const cplx<T> s02 = cadd(a0, a2), d02 = csub(a0, a2);
  const cplx<T> s13 = cadd(a1, a3), d13 = cmul_i<DIR>(csub(a1, a3));
  a0 = cadd(s02, s13);
  a1 = cadd(d02, d13);
  a2 = csub(s02, s13);
  a3 = csub(d02, d13);
}

// multiply by exp(DIR * 2*pi*i * m / 16), m a compile-time constant
template <int DIR, int M16, typename T>
MLX_HD cplx<T> cmul_w16(const cplx<T> a) {
  constexpr int m = ((M16 % 16) + 16) % 16;
  if constexpr (m == 0) {
    return a;
  } else if constexpr (m == 4) {
    return cmul_i<DIR>(a);
  } else if constexpr (m == 8) {
    return cplx<T>{-a.x, -a.y};
  } else if constexpr (m == 12) {
    return cmul_i<-DIR>(a);
  } else if constexpr (m % 2 == 0) {  // odd multiples of pi/4
    constexpr T h = T(0.70710678118654752440084436210485L);
    // (a.x + i a.y) * (c + i s) with |c| = |s| = h
    constexpr int q = m / 2;  // 1,3,5,7
    constexpr T c = (q == 1 || q == 7) ? h : -h;
    constexpr T s0 = (q == 1 || q == 3) ? h : -h;  // sin(2 pi m/16) sign
    constexpr T s = DIR > 0 ? s0 : -s0;
    return cplx<T>{a.x * c - a.y * s, a.x * s + a.y * c};
  } else {
    constexpr T c1 = T(0.92387953251128675612818318939679L);  // cos(pi/8)
    constexpr T s1 = T(0.38268343236508977172845998403040L);  // sin(pi/8)
    // cos/sin(2 pi m / 16) for odd m
    constexpr T c = (m == 1 || m == 15) ? c1 : (m == 3 || m == 13) ? s1 : (m == 5 || m == 11) ? -s1 : -c1;
    constexpr T s0 = (m == 1 || m == 7) ? s1 : (m == 3 || m == 5) ? c1 : (m == 9 || m == 15) ? -s1 : -c1;
    constexpr T s = DIR > 0 ? s0 : -s0;
    return cplx<T>{a.x * c - a.y * s, a.x * s + a.y * c};
  }
}

// The same rotation with the order of operations pinned (one rounded product, one fused multiply-add per
// component): used where a twiddle is DERIVED from another one in two different kernels that must agree bit
// for bit -- left to the compiler, which of the two products gets fused may differ between inlined copies.
template <int DIR, int M16>
MLX_HD cplx<double> rot16_pinned(const cplx<double> a) {
  constexpr int m = ((M16 % 16) + 16) % 16;
  if constexpr (m % 4 == 0) {
    return cmul_w16<DIR, M16>(a);  // exact: no rounding at all
  } else {
    constexpr double h = 0.70710678118654752440084436210485;
    constexpr double c1 = 0.92387953251128675612818318939679, s1 = 0.38268343236508977172845998403040;
    constexpr double c = (m % 2 == 0) ? ((m == 2 || m == 14) ? h : -h)
                                      : ((m == 1 || m == 15) ? c1 : (m == 3 || m == 13) ? s1 : (m == 5 || m == 11) ? -s1 : -c1);
    constexpr double s0 = (m % 2 == 0) ? ((m == 2 || m == 6) ? h : -h)
                                       : ((m == 1 || m == 7) ? s1 : (m == 3 || m == 5) ? c1 : (m == 9 || m == 15) ? -s1 : -c1);
    constexpr double s = DIR > 0 ? s0 : -s0;
#ifdef __CUDA_ARCH__
    return cplx<double>{fma(a.x, c, -__dmul_rn(a.y, s)), fma(a.x, s, __dmul_rn(a.y, c))};
#else
    return cplx<double>{a.x * c - a.y * s, a.x * s + a.y * c};
#endif
  }
}

template <int DIR, typename T>
MLX_HD void dft8(cplx<T> (&v)[8]) {
  // n = 2a + b, k = c + 4d
  dft4<DIR>(v[0], v[2], v[4], v[6]);  // y_0[c] -> v[0],v[2],v[4],v[6]
  dft4<DIR>(v[1], v[3], v[5], v[7]);  // y_1[c] -> v[1],v[3],v[5],v[7]
  const cplx<T> y0[4] = {v[0], v[2], v[4], v[6]};
  cplx<T> y1[4] = {v[1], cmul_w16<DIR, 2>(v[3]), cmul_w16<DIR, 4>(v[5]), cmul_w16<DIR, 6>(v[7])};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    v[c] = cadd(y0[c], y1[c]);
    v[c + 4] = csub(y0[c], y1[c]);
  }
}

template <int DIR, typename T>
MLX_HD void dft16(cplx<T> (&v)[16]) {
  // n = 4a + b, k = c + 4d:  y_b[c] = DFT4_a x[4a+b];  y_b[c] *= W16^{bc};  X[c+4d] = DFT4_b y_b[c]
  dft4<DIR>(v[0], v[4], v[8], v[12]);   // b = 0: y_0[c] in v[4c]
  dft4<DIR>(v[1], v[5], v[9], v[13]);   // b = 1: y_1[c] in v[4c+1]
  dft4<DIR>(v[2], v[6], v[10], v[14]);  // b = 2
  dft4<DIR>(v[3], v[7], v[11], v[15]);  // b = 3
  // twiddles W16^{b c}: element v[4c + b]
  v[5] = cmul_w16<DIR, 1>(v[5]);
  v[6] = cmul_w16<DIR, 2>(v[6]);
  v[7] = cmul_w16<DIR, 3>(v[7]);
  v[9] = cmul_w16<DIR, 2>(v[9]);
  v[10] = cmul_w16<DIR, 4>(v[10]);
  v[11] = cmul_w16<DIR, 6>(v[11]);
  v[13] = cmul_w16<DIR, 3>(v[13]);
  v[14] = cmul_w16<DIR, 6>(v[14]);
  v[15] = cmul_w16<DIR, 9>(v[15]);
  // for each c: DFT4 over b of v[4c + b] -> X[c + 4d] ; result d lands in v[4c + d]
  dft4<DIR>(v[0], v[1], v[2], v[3]);
  dft4<DIR>(v[4], v[5], v[6], v[7]);
  dft4<DIR>(v[8], v[9], v[10], v[11]);
  dft4<DIR>(v[12], v[13], v[14], v[15]);
  // now v[4c + d] = X[c + 4d]: transpose the 4x4 index to natural order
  cplx<T> t;
#define MLX_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
  MLX_SWAP(1, 4) MLX_SWAP(2, 8) MLX_SWAP(3, 12) MLX_SWAP(6, 9) MLX_SWAP(7, 13) MLX_SWAP(11, 14)
#undef MLX_SWAP
}

template <int R, int DIR, typename T>
MLX_HD void dft_r(cplx<T> (&v)[R]) {
  if constexpr (R == 2) dft2<DIR>(v[0], v[1]);
  else if constexpr (R == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 8) dft8<DIR>(v);
  else dft16<DIR>(v);
}

// p[r-1] = w^r, r = 1..15, by a product tree of depth <= 4 (keeps the float twiddle error at a few ulp)
template <typename T>
MLX_HD void power_tree16(cplx<T> (&p)[15], const cplx<T> w) {
  const cplx<T> w2 = cmul(w, w), w3 = cmul(w2, w), w4 = cmul(w2, w2);
  const cplx<T> w5 = cmul(w4, w), w6 = cmul(w4, w2), w7 = cmul(w4, w3), w8 = cmul(w4, w4);
  p[0] = w; p[1] = w2; p[2] = w3; p[3] = w4; p[4] = w5; p[5] = w6; p[6] = w7; p[7] = w8;
  p[8] = cmul(w8, w); p[9] = cmul(w8, w2); p[10] = cmul(w8, w3); p[11] = cmul(w8, w4);
  p[12] = cmul(w8, w5); p[13] = cmul(w8, w6); p[14] = cmul(w8, w7);
}

// v[r] *= w^r, r = 1..R-1
template <int R, typename T>
MLX_HD void twiddle_powers(cplx<T> (&v)[R], const cplx<T> w) {
  if constexpr (R == 2) {
    v[1] = cmul(v[1], w);
  } else if constexpr (R == 4) {
    const cplx<T> w2 = cmul(w, w);
    v[1] = cmul(v[1], w);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], cmul(w2, w));
  } else if constexpr (sizeof(T) == 8 && !MLX_FFT_TREE64) {
    // double: linear recurrence is accurate to ~R ulp(double), far more than needed
    cplx<T> p = w;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      v[r] = cmul(v[r], p);
      if (r + 1 < R) p = cmul(p, w);
    }
  } else if constexpr (R == 16) {
    cplx<T> p[15];
    power_tree16(p, w);
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], p[r - 1]);
  } else {
    // float, radix 8: the first half of the same tree
    const cplx<T> w2 = cmul(w, w), w3 = cmul(w2, w), w4 = cmul(w2, w2);
    const cplx<T> w5 = cmul(w4, w), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
    v[1] = cmul(v[1], w);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], w5);
    v[6] = cmul(v[6], w6);
    v[7] = cmul(v[7], w7);
  }
}

// v[r] *= tab[(r - 1) * 16 + k], r = 1..15: the fifteen powers of the second stage's twiddle exp(DIR 2 pi i k / 256)
// from a 240-entry table (k < 16) instead of a chain of fourteen complex products.  Used by the double-precision
// analysis transform, where the chain is 56 dependent FP64 instructions per thread and frame; the table values are
// correctly rounded, the chain drifts by a few ulp.
template <typename T>
MLX_HD void twiddle_table16(cplx<T> (&v)[16], const cplx<T>* tab, int k) {
#pragma unroll
  for (int r = 1; r < 16; ++r) {
#ifdef __CUDA_ARCH__
    cplx<T> w;
    if constexpr (sizeof(T) == 8) {
      const double2 q = __ldg(reinterpret_cast<const double2*>(tab + (r - 1) * 16 + k));
      w = cplx<T>{(T)q.x, (T)q.y};
    } else {
      const float2 q = __ldg(reinterpret_cast<const float2*>(tab + (r - 1) * 16 + k));
      w = cplx<T>{(T)q.x, (T)q.y};
    }
#else
    const cplx<T> w = tab[(r - 1) * 16 + k];
#endif
    v[r] = cmul(v[r], w);
  }
}

// ---------------------------------------------------------------- 32 points per thread (NC = 1024 = 32 x 32)
// cos / sin as compile-time constants (Taylor series in double; |x| <= 2 pi)
MLX_HDC double fft_ccos(double x) {
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 28; ++i) {
    term *= -x * x / ((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
MLX_HDC double fft_csin(double x) {
  double term = x, sum = x;
  for (int i = 1; i < 28; ++i) {
    term *= -x * x / ((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
struct W32Tab {  // exp(+2 pi i k / 32), k < 16
  float c[16], s[16];
};
MLX_HDC W32Tab make_w32() {
  W32Tab r{};
  for (int k = 0; k < 16; ++k) {
    r.c[k] = (float)fft_ccos(6.283185307179586476925286766559 * k / 32.0);
    r.s[k] = (float)fft_csin(6.283185307179586476925286766559 * k / 32.0);
  }
  return r;
}
// 32-point DFT, natural order in and out: two interleaved 16-point transforms and one radix-2 level
template <int DIR>
MLX_HD void dft32(cplx<float> (&v)[32]) {
  using C = cplx<float>;
  C e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    e[i] = v[2 * i];
    o[i] = v[2 * i + 1];
  }
  dft16<DIR>(e);
  dft16<DIR>(o);
  constexpr W32Tab w = make_w32();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    C tw;  // o[k] * exp(DIR 2 pi i k / 32)
    if (k == 0) {
      tw = o[k];
    } else if (k == 8) {
      tw = cmul_i<DIR>(o[k]);
    } else {
      const float sn = DIR > 0 ? w.s[k] : -w.s[k];
      tw = C{o[k].x * w.c[k] - o[k].y * sn, o[k].x * sn + o[k].y * w.c[k]};
    }
    v[k] = cadd(e[k], tw);
    v[k + 16] = csub(e[k], tw);
  }
}
// one pad element per 32: the stride-33 stores of the first radix-32 stage and the unit-stride loads of the
// second are conflict-free
MLX_HDC int pad32(int i) { return i + (i >> 5); }

// The 1024-point transform of one frame by 32 threads (one warp), 32 points each: x[r] = in[t + 32 r] on entry and
// out[t + 32 r] on return.  p1[r - 1] = exp(DIR 2 pi i t r / 1024), r = 1..31 (kept in registers by the caller).
// `sync()` orders the group's shared-memory accesses (__syncwarp on the device; the host emulation runs the two halves
// as separate passes over the threads).
struct Fft32x32 {
  using C = cplx<float>;
  static constexpr int NC = 1024, BUF = NC + NC / 32;
  template <int DIR>
  static MLX_HD void stage0(C (&x)[32], C* buf, int t) {  // butterfly t; output r goes to position 32 t + r
    dft32<DIR>(x);
    C* p = buf + 33 * t;  // pad32(32 t + r) = 33 t + r
#pragma unroll
    for (int r = 0; r < 32; ++r) p[r] = x[r];
  }
  template <int DIR>
  static MLX_HD void stage1(C (&x)[32], const C* buf, int t, const C (&p1)[31]) {
    const C* p = buf + t;  // pad32(t + 32 r) = t + 33 r
#pragma unroll
    for (int r = 0; r < 32; ++r) x[r] = p[33 * r];
#pragma unroll
    for (int r = 1; r < 32; ++r) x[r] = cmul(x[r], p1[r - 1]);
    dft32<DIR>(x);  // output r is element t + 32 r
  }
  static MLX_HD void store(const C (&x)[32], C* buf, int t) {  // natural order, padded
    C* p = buf + t;
#pragma unroll
    for (int r = 0; r < 32; ++r) p[33 * r] = x[r];
  }
};

// ---------------------------------------------------------------- plan
template <int NC>
struct FftPlan {
  static_assert(NC >= 256 && (NC & (NC - 1)) == 0, "NC must be a power of two >= 256");
  static MLX_HDC int log2nc() {
    int l = 0;
    while ((1 << l) < NC) ++l;
    return l;
  }
  static constexpr int LOG = log2nc();
  static constexpr int A = LOG / 4;  // radix-16 stages
  static constexpr int B = LOG % 4;  // last stage radix 2^B (absent when B == 0)
  static constexpr int NSTAGES = A + (B ? 1 : 0);
  static constexpr int TPF = NC / 16;              // threads per frame
  static constexpr int BUF = NC + NC / 16;         // padded complex elements per frame buffer
  static MLX_HDC int radix(int s) { return s < A ? 16 : (1 << B); }
  static MLX_HDC int ns(int s) { return s < A ? (1 << (4 * s)) : (1 << (4 * A)); }
  // twiddle registers: one per butterfly per stage >= 1
  static MLX_HDC int nw() {
    int n = 0;
    for (int s = 1; s < NSTAGES; ++s) n += 16 / radix(s);
    return n;
  }
  static constexpr int NW = nw() > 0 ? nw() : 1;
  static MLX_HDC int woff(int s) {
    int n = 0;
    for (int q = 1; q < s; ++q) n += 16 / radix(q);
    return n;
  }
};

// Per-thread twiddles (constant across frames): w[woff(s) + b] = exp(DIR*2*pi*i*k/(NS*R)),
// k = (t + b*TPF) mod NS.  `table[m] = exp(-2*pi*i*m/NC)` (forward), m in [0, NC).
// PRE1: also keep the 15 powers of the stage-1 twiddle (radix-16 second stage, one butterfly per
// thread) in registers instead of rebuilding them per transform -- the same product tree, so the
// results are bit-identical; costs 30 registers, saves 14 complex products per transform.
// TAB1: the powers of the stage-1 twiddle come from `tab1` (twiddle_table16; forward direction only: the table holds
// exp(-2 pi i k r / 256)), the stage-1 twiddle register is not kept.
template <typename T, int NC, int DIR, bool PRE1 = false, bool TAB1 = false>
struct FftTwiddles {
  using P = FftPlan<NC>;
  static constexpr bool kPre1 = PRE1 && P::NSTAGES >= 2 && P::radix(1) == 16 && sizeof(T) == 4;
  static constexpr bool kTab1 = TAB1 && P::NSTAGES >= 2 && P::radix(1) == 16 && DIR < 0;
  cplx<T> w[P::NW];
  cplx<T> p1[kPre1 ? 15 : 1];
  const cplx<T>* tab1 = nullptr;
  MLX_HD void init(int t, const cplx<T>* table) {
#pragma unroll
    for (int s = 1; s < P::NSTAGES; ++s) {
      const int R = P::radix(s), NS = P::ns(s), BPT = 16 / R;
      if (kTab1 && s == 1) continue;  // read from the table
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        const int j = t + b * P::TPF;
        const int k = j & (NS - 1);
        cplx<T> v = table[k * (NC / (NS * R))];
        if (DIR > 0) v.y = -v.y;
        w[P::woff(s) + b] = v;
      }
    }
    if constexpr (kPre1) {
      cplx<T> p[15];
      power_tree16(p, w[P::woff(1)]);
#pragma unroll
      for (int r = 0; r < 15; ++r) p1[r] = p[r];
    }
  }
};

template <typename T, int NC, int DIR>
struct Fft {
  using P = FftPlan<NC>;
  using C = cplx<T>;
  static constexpr int TPF = P::TPF;

  // Padded address of element (t + m*TPF): TPF is a multiple of 16, so the padding term splits
  // exactly, pad(t + m*TPF) = pad(t) + m*(TPF + TPF/16) -- one base pointer per thread and
  // compile-time offsets instead of address arithmetic per access.
  static constexpr int SLOT_STRIDE = TPF + TPF / 16;
  static_assert(TPF % 16 == 0, "slot stride needs TPF to be a multiple of 16");

  // load slot m <- buf[t + m*TPF]
  static MLX_HD void load(C (&x)[16], const C* buf, int t) {
    const C* p = buf + fft_pad(t);
#pragma unroll
    for (int m = 0; m < 16; ++m) x[m] = p[m * SLOT_STRIDE];
  }
  // store slot m -> buf[t + m*TPF]
  static MLX_HD void store(const C (&x)[16], C* buf, int t) {
    C* p = buf + fft_pad(t);
#pragma unroll
    for (int m = 0; m < 16; ++m) p[m * SLOT_STRIDE] = x[m];
  }

  // butterflies of stage S on the register slots; non-last stages scatter to buf, the last stage
  // leaves natural-order results in the slots.
  template <int S, class TW>
  static MLX_HD void compute(C (&x)[16], C* buf, int t, const TW& tw) {
    constexpr int R = P::radix(S), NS = P::ns(S), BPT = 16 / R;
    constexpr bool LAST = (S == P::NSTAGES - 1);
#pragma unroll
    for (int b = 0; b < BPT; ++b) {
      C v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = x[b + r * BPT];
      if constexpr (S == 1 && TW::kTab1) {
        twiddle_table16(v, tw.tab1, (t + b * TPF) & (NS - 1));
      } else if constexpr (S == 1 && TW::kPre1) {
#pragma unroll
        for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw.p1[r - 1]);
      } else if constexpr (S > 0 && LAST && MLX_FFT_DERIVE_LAST && BPT > 1 && sizeof(T) == 8) {
        // last stage: j = t + b*TPF < NS, so w_b = exp(DIR 2 pi i (t + b*TPF) / NC) = w_0 * exp(DIR 2 pi i b / 16)
        const C w0 = tw.w[P::woff(S)];
        C wb = w0;
        switch (b) {
          case 1: wb = rot16_pinned<DIR, 1>(w0); break;
          case 2: wb = rot16_pinned<DIR, 2>(w0); break;
          case 3: wb = rot16_pinned<DIR, 3>(w0); break;
          case 4: wb = rot16_pinned<DIR, 4>(w0); break;
          case 5: wb = rot16_pinned<DIR, 5>(w0); break;
          case 6: wb = rot16_pinned<DIR, 6>(w0); break;
          case 7: wb = rot16_pinned<DIR, 7>(w0); break;
          default: break;
        }
        twiddle_powers<R>(v, wb);
      } else if constexpr (S > 0) {
        twiddle_powers<R>(v, tw.w[P::woff(S) + b]);
      }
      dft_r<R, DIR>(v);
      if constexpr (LAST) {
#pragma unroll
        for (int r = 0; r < R; ++r) x[b + r * BPT] = v[r];
      } else {
        const int j = t + b * TPF;
        const int k = j & (NS - 1);
        const int j0 = (j - k) * R + k;
        // pad(j0 + r*NS): for NS = 1 (first stage) j0 = 16 j and r < 16 stay inside one padding
        // block; for NS a multiple of 16 the padding term splits exactly.
        C* p = buf + fft_pad(j0);
        constexpr int RS = (NS % 16 == 0) ? NS + NS / 16 : NS;
        static_assert(NS % 16 == 0 || (NS == 1 && R == 16), "unsupported stage geometry");
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * RS] = v[r];
      }
    }
  }

#ifdef __CUDACC__
  // Full transform for one thread of the group.  in: x[m] = in[t + m*TPF]; out: x[m] = out[t + m*TPF].
  // `buf` must not be in use by the group when run() is entered.  Bar::sync() synchronises the group.
  template <int S, class TW, class Bar>
  static __device__ __forceinline__ void run_from(C (&x)[16], C* buf, int t, const TW& tw, Bar& bar) {
    if constexpr (S < P::NSTAGES) {
      if constexpr (S > 0) {
        bar.sync();  // stage S-1 stores visible
        load(x, buf, t);
        if constexpr (S < P::NSTAGES - 1) bar.sync();  // all loads done before anyone overwrites
      }
      compute<S>(x, buf, t, tw);
      run_from<S + 1>(x, buf, t, tw, bar);
    }
  }
  template <class TW, class Bar>
  static __device__ __forceinline__ void run(C (&x)[16], C* buf, int t, const TW& tw, Bar& bar) {
    run_from<0>(x, buf, t, tw, bar);
  }
#endif
};

}  // namespace mlx
