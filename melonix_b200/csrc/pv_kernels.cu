// melonix_b200/csrc/pv_kernels.cu -- phase-vocoder kernels for sm_100a (PV-spec v1).
//
// NOT IN REFERENCE: melonix has no phase vocoder; these kernels implement the
// north-star's analysis / pitch-detect / bin-shift / resynthesis path.  Frame geometry follows the
// reference's Spec jobs (spec.cpp:47, spec-cache.cpp:63-65).
//
//   pv_analyze_kernel  K_A  one CTA = one chunk of consecutive frames of one track, G frames per
//                           batch (256 threads x 2 CTAs per SM for fftN <= 2048, 512 x 1 beyond).
//                           TMA bulk load of the batch's sample tile -> Hann window ->
//                           FP64 real FFT (N/2-point complex Stockham in shared memory) ->
//                           d = arg(X_f conj(X_{f-1}) (-i)^k) -> magnitude, peak bin, f0 ->
//                           bin shift with per-bin launch constants -> exact uint32 phase
//                           increments, chunk-local scan.
//   pv_scan_kernel          exclusive scan of the chunk totals per (track, bin).
//   pv_synth_kernel    K_S  theta = prefix + local sum -> Y = smag e^{i theta} (MUFU sin/cos of the
//                           exact integer phase) -> FP32 inverse real FFT -> synthesis window ->
//                           atomics-free overlap-add (fixed ascending-frame summation order, three
//                           pending hops in registers) -> float2 stores.
//
// Why FP64 in K_A: the wrapped phase difference has a cut at +-pi; a frame whose FP32 spectrum
// lands on the other side of the cut than the oracle's shifts that bin's accumulated phase by
// frac(rate) turns for the rest of the track.  B200 runs FP64 at half the FP32 rate, which makes
// the analysis transform robust for ~2x its FP32 cost.  Everything after the cut decision is FP32 /
// exact uint32.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "fft.cuh"
#include "kernels.h"
#include "pv_analysis.cuh"
#include "pv_shift.cuh"
#include "tma.cuh"
#include "pv_common.cuh"
#include "spec_frame.cuh"  // spec_rot32 / spec_mul_here: twiddles derived by compile-time rotations

#ifndef MLX_KA_DERIVE_WPAIR
#define MLX_KA_DERIVE_WPAIR 1  // one pair-split twiddle per thread, the others by constant rotation (4 registers)
#endif
#ifndef MLX_KA_TOTC_DIRECT
#define MLX_KA_TOTC_DIRECT 1  // FAST instantiation: the phase at the wave end goes to memory at frame we - 1 instead of
                              // riding along in QB registers (196 -> 96 bytes of spills; 13.8 -> 13.2 ms)
#endif
#ifndef MLX_UNROLL_PAIR
#define MLX_UNROLL_PAIR 4
#endif
#ifndef MLX_UNROLL_GATHER
#define MLX_UNROLL_GATHER 2
#endif
#ifndef MLX_GATHER_V2
#define MLX_GATHER_V2 1  // constant-rate bin shift with frame-invariant constants (bit-identical to v1)
#endif
#ifndef MLX_KS_MINB3_MAXN
#define MLX_KS_MINB3_MAXN 2048  // up to this fftN: three synthesis CTAs per SM (80 registers), two beyond
#endif
#ifndef MLX_KS_PRE1
#define MLX_KS_PRE1 0      // keep the 15 stage-1 twiddle powers of the inverse FFT in registers
#endif
namespace mlx {

constexpr int kUnrollPair = MLX_UNROLL_PAIR;      // frames of the pair phase processed together (ILP)
constexpr int kUnrollGather = MLX_UNROLL_GATHER;  // frames of the gather phase processed together

// ------------------------------------------------------------------------------------------------
// scalar helpers shared by K_A / K_S

// trunc(float(k) * r) with a plain float multiply, exactly as the spec (A.5) and the oracle do.
__device__ __forceinline__ int shift_bin(int k, float r) { return (int)truncf(__fmul_rn((float)k, r)); }

// Slow path for per-frame rates: K_j = { k in [0, NC] : shift_bin(k, r) == j } searched on the
// device (klo > khi when empty) plus the base increment.  The constant-rate path reads tables.
__device__ __noinline__ void gather_entry_slow(int j, float r, int NC, uint32_t& kk) {
  int k = (int)(__fdividef((float)j, r)) - 1;
  k = max(0, min(k, NC));
  while (k <= NC && shift_bin(k, r) < j) ++k;
  while (k > 0 && shift_bin(k - 1, r) >= j) --k;
  if (k > NC || shift_bin(k, r) != j) {
    kk = 1u;  // klo = 1, khi = 0
    return;
  }
  int khi = k;
  while (khi + 1 <= NC && shift_bin(khi + 1, r) == j) ++khi;
  kk = (uint32_t)k | ((uint32_t)khi << 16);
}

// Bin shift + exact phase increment of one output bin j for one frame (PV-spec A.5/A.6).
// `zb` holds the frame's (mag, d) records.  Returns the shifted magnitude; inc = uint32 increment.
template <int NC, int BUF>
__device__ __forceinline__ float shift_one_bin(const cplx<double>* zb, int j, uint32_t kk, uint32_t r_fix,
                                               uint32_t& inc) {
  // records of the pair (k, NC - k) share the 16-byte slot of k <= NC/2 (pv_shift.cuh: rec_offset)
  auto magd = [&](int k) {
    return *reinterpret_cast<const MagD*>(reinterpret_cast<const unsigned char*>(zb) + rec_offset(k, NC, true));
  };
  const int klo = (int)(kk & 0xffffu), khi = (int)(kk >> 16);
  const bool any = klo <= khi;  // K_j non-empty (at most one bin when rate >= 1)
  const int kh = any ? khi : 0;
  const MagD mh = magd(kh);  // the bin whose frequency the output bin inherits
  float smag = any ? fabsf(klo == khi ? mh.mag : magd(klo).mag) : 0.f;
  for (int k = klo + 1; k <= khi; ++k) smag += fabsf(magd(k).mag);  // only when rate < 1
  // frac(rate * nu / 4) * 2^32 with nu / 4 = (khi * 2^30 + d) / 2^32 turns, d the signed phase advance
  // (+-2^32 when the cut decision says so): one exact product mod 2^64, rounded once -- the same
  // single rounding per frame as the oracle's llrint.  32-bit pieces:
  const int d32 = mh.d;
  int dhi = d32 >> 31;
  dhi += (__float_as_uint(mh.mag) >> 31) ? ((d32 < 0) ? 1 : -1) : 0;
  const uint32_t k30 = (uint32_t)kh << 30;
  const uint32_t nlo = k30 + (uint32_t)d32;
  const uint32_t nhi = (uint32_t)(kh >> 2) + (uint32_t)dhi + (nlo < k30 ? 1u : 0u);
  const unsigned long long prod =
      (unsigned long long)r_fix * nlo + ((unsigned long long)(r_fix * nhi) << 32) + (1ULL << 25);
  inc = any ? (uint32_t)(prod >> 26) : (((uint32_t)j & 3u) << 30);  // empty K_j: s_nu = j -> frac(j / 4)
  return smag;
}

// one frame of one output bin whose K_j has at most one element (rate >= 1); constants: pv_shift.cuh
__device__ __forceinline__ float shift_one_bin_v2(const cplx<double>* zb, const ShiftConstA& c, int r_fix,
                                                  uint32_t& inc) {
  const MagD mh = *reinterpret_cast<const MagD*>(reinterpret_cast<const unsigned char*>(zb) + c.off);
  const uint32_t mb = __float_as_uint(mh.mag);
  inc = shift_inc_a(c.A, mh.d, mb, r_fix);
  return __uint_as_float(mb & 0x7fffffffu);
}

#ifndef MLX_SINCOS_MUFU
#define MLX_SINCOS_MUFU 1
#endif
#if MLX_SINCOS_MUFU
// sin/cos of 2*pi*acc/2^32 on the special-function unit: the signed phase in [-pi, pi) goes through
// sin.approx / cos.approx (SASS: I2FP, FMUL, FMUL.RZ, MUFU.SIN, MUFU.COS -- 5 instructions instead of
// the 23 of the polynomial version below).  Absolute error <= ~6e-7 (MUFU 2^-21.4 plus the float
// rounding of the 32-bit phase); it does not accumulate (the phase itself is the exact integer) and
// moves the output by ~1e-7 RMS, three orders of magnitude inside the 1e-4 budget.
__device__ __forceinline__ void sincos_turns(uint32_t acc, float& s, float& c) {
  const float x = (float)(int)acc * 1.4629180792671596e-09f;  // 2 pi / 2^32
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
}
#else
// sin/cos of 2*pi*acc/2^32: the quadrant comes from the top bits (exact range reduction for free),
// the remainder phi in [-pi/4, pi/4) goes through degree-7/8 polynomials (6e-8 max error in float).
__device__ __forceinline__ void sincos_turns(uint32_t acc, float& s, float& c) {
  const uint32_t k = (acc + 0x20000000u) >> 30;
  const int f = (int)(acc - (k << 30));  // [-2^29, 2^29)
  // (f >> 7) as float without an I2F: exponent trick, exact for |.| < 2^22
  const float fq = __int_as_float(0x4B400000 + (f >> 7)) - 12582912.0f;
  const float phi = fq * 1.8725351414619643e-07f;  // 2 pi * 2^7 / 2^32
  const float s2 = phi * phi;
  float sp = -1.950396254e-04f;
  sp = fmaf(sp, s2, 8.332036436e-03f);
  sp = fmaf(sp, s2, -1.666665077e-01f);
  sp = fmaf(sp, s2, 1.0f);
  sp *= phi;
  float cp = 2.437988041e-05f;
  cp = fmaf(cp, s2, -1.388661913e-03f);
  cp = fmaf(cp, s2, 4.166661575e-02f);
  cp = fmaf(cp, s2, -0.5f);
  cp = fmaf(cp, s2, 1.0f);
  const bool sw = k & 1u;
  float ss = sw ? cp : sp, cc = sw ? sp : cp;
  s = (k & 2u) ? -ss : ss;
  c = ((k + 1u) & 2u) ? -cc : cc;
}
#endif

// ------------------------------------------------------------------------------------------------
// K_A
// FAST: the launch has one constant ratio >= 1 for every track (the host checks): the per-frame-rate and
// multi-bin-gather code, and the registers it keeps alive, are compiled out.
template <int N, int G, bool FAST>
__global__ void __launch_bounds__(PvCfg<N, G>::THREADS, PvG<N>::ka_ctas)
pv_analyze_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = PvCfg<N, G>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, H = Cfg::H, NB = Cfg::NB, NBP = Cfg::NBP;
  constexpr int THREADS = Cfg::THREADS, BUF = Cfg::BUFS, TILE = Cfg::TILE, QP = Cfg::QP, QB = Cfg::QB;
  constexpr bool WD = Cfg::WIN_D;
  using C = cplx<double>;
  using F = Fft<double, NC, -1>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                        // [G][BUF]  Z, then (mag, d) per bin
  double* s_win = reinterpret_cast<double*>(buf + G * BUF);       // [N] when WD
  float* tile = reinterpret_cast<float*>(s_win + (WD ? N : 0));   // [2][TILE]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tile + 2 * TILE);  // [2]
  constexpr int ZSLOT = BUF - 1;  // (mag, d) = (0, 0): what an empty K_j reads
  auto rec = [](const C* zb, int k) {  // (mag, d) record of input bin k (pv_shift.cuh: rec_offset)
    return *reinterpret_cast<const MagD*>(reinterpret_cast<const unsigned char*>(zb) + rec_offset(k, NC, true));
  };

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  const PvTrack tr = tracks[blockIdx.y];
  const long long lim = min(wv.we + 3, tr.F);  // analysis runs three frames past the owned window
  const long long a = wv.wb + (long long)blockIdx.x * wv.CA;
  const size_t trow = ((size_t)blockIdx.y * wv.nchunksA + blockIdx.x) * NBP;
  if (a >= lim) {  // chunk past the end of this track: contributes nothing to the scan
    for (int j = tid; j < NB; j += THREADS) {
      sc.tot[trow + j] = 0u;
      sc.totc[trow + j] = 0u;
    }
    return;
  }
  const long long b = min(a + (long long)wv.CA, lim);
  const int nbatch = (int)((b - a + 1 + G - 1) / G);  // frames a-1 .. b-1

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
  }
  if constexpr (WD) {
    // staged at HALF scale: the pair split X = E + W O works on Z / 2 and needs no halving of its own (a power
    // of two: every product and sum of the transform scales exactly, the results are bit-identical)
    for (int i = tid; i < N; i += THREADS) s_win[i] = 0.5 * tb.win_d[i];
  }
  constexpr double kSplit = WD ? 1.0 : 0.5;  // what the pair split still has to apply
  __syncthreads();  // mbarriers initialised, window staged
  if (tid == 0) {   // first tile: samples [(a-1-3)H, (a-1+G)H) of the zero-padded track
    mbar_expect_tx(mbar, TILE * sizeof(float));
    tma_load_1d(tile, tr.x + (a - 4) * H, TILE * sizeof(float), mbar);
  }

#ifndef MLX_KA_TAB1
#define MLX_KA_TAB1 0  // 1: stage-1 twiddle powers of the FP64 transform from a 240-entry table (L1) instead of a chain of
                       // fourteen dependent complex products per thread and frame -- measured: 13.83 ms against 13.20
                       // (fifteen 16-byte L1 loads per thread cost more than the 56 FP64 instructions they replace)
#endif
  FftTwiddles<double, NC, -1, false, MLX_KA_TAB1 != 0> twd;
  twd.init(t, tb.tw_d);
  twd.tab1 = tb.tw1_d;
  // pair slots: bins k and NC-k for k = 1 + tid + q*THREADS <= NC/2; thread 0 also owns (0, NC)
  C wpair[QP];  // exp(-2 pi i k / N) of this thread's pairs
#pragma unroll
  for (int q = 0; q < QP; ++q) {
    const int k = 1 + tid + q * THREADS;
    wpair[q] = tb.twr_d[k <= NC / 2 ? k : 0];
  }
  // previous frame per bin: X (double, for the cut decision), integer phase, magnitude.
  // Frame -1 has phi = 0 <=> X = 1.
  C pk[QP], pm[QP];
  uint32_t ppk[QP], ppm[QP];
  float pmk[QP], pmm[QP];
#pragma unroll
  for (int q = 0; q < QP; ++q) {
    pk[q] = pm[q] = C{1.0, 0.0};
    ppk[q] = ppm[q] = 0u;
    pmk[q] = pmm[q] = 1.f;
  }
  uint32_t pp0 = 0u, ppn = 0u;
  float pm0 = 1.f, pmn = 1.f;
  uint32_t lacc[QB], totc[QB];  // chunk-local phase sum; its value at the last frame < we
#pragma unroll
  for (int q = 0; q < QB; ++q) lacc[q] = totc[q] = 0u;
  uint32_t lacc_nyq = 0u, totc_nyq = 0u;  // bin NC, kept by every lane of the last warp
  const bool per_frame_rate = FAST ? false : (tr.rate_pf != nullptr);
  // constant-rate path: the bin-shift table entries of this thread's bins are frame-invariant
  uint32_t gkq[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    const int j = tid + q * THREADS;
    gkq[q] = (!per_frame_rate && j < NC) ? __ldg(wv.gk + j) : 1u;
  }
  const uint32_t gk_nyq = per_frame_rate ? 1u : __ldg(wv.gk + NC);
#if MLX_GATHER_V2
  // rate >= 1: every K_j holds at most one bin and the shift of a bin is a handful of integer ops
  const bool fast_shift = FAST ? true : (!per_frame_rate && wv.rate >= 1.0f);
  ShiftConstA scq[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q)
    scq[q] = make_shift_const_a(tid + q * THREADS, gkq[q], (uint32_t)wv.r_fix, NC, true, 16u * ZSLOT);
  for (int gg = tid; gg < G; gg += THREADS) buf[gg * BUF + ZSLOT] = C{0.0, 0.0};  // visible after the first barrier
#endif
  const GroupBar<TPF> bar = make_group_bar<TPF>(g, tid);
  const size_t row0 = (size_t)blockIdx.y * wv.rows;

  for (int bi = 0; bi < nbatch; ++bi) {
    const long long f_first = a - 1 + (long long)bi * G;
    float* cur = tile + (bi & 1) * TILE;
    // ---- TMA: prefetch the next batch's tile, then wait for this one
    if (tid == 0 && bi + 1 < nbatch) {
      fence_proxy_async();  // the tile's generic-proxy reads (two batches ago, ordered by __syncthreads) before its refill
      mbar_expect_tx(mbar + ((bi + 1) & 1), TILE * sizeof(float));
      tma_load_1d(tile + ((bi + 1) & 1) * TILE, tr.x + (f_first + G - 3) * H, TILE * sizeof(float),
                  mbar + ((bi + 1) & 1));
    }
    mbar_wait(mbar + (bi & 1), (bi >> 1) & 1);

    // ---- forward FP64 real FFT of frame f_first + g (N/2-point complex on even/odd samples)
    const long long fg = f_first + g;
    if (fg >= 0 && fg < b) {
      C x[16];
      const float* src = cur + g * H;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int i = t + m * TPF;
        const float2 s2 = *reinterpret_cast<const float2*>(src + 2 * i);
        if constexpr (WD) {
          const double2 w2 = *reinterpret_cast<const double2*>(s_win + 2 * i);
          x[m] = C{w2.x * (double)s2.x, w2.y * (double)s2.y};
        } else {
          const float2 w2 = __ldg(reinterpret_cast<const float2*>(tb.win + 2 * i));
          x[m] = C{(double)w2.x * (double)s2.x, (double)w2.y * (double)s2.y};
        }
      }
      F::run(x, buf + g * BUF, t, twd, bar);
      F::store(x, buf + g * BUF, t);  // same thread-private slots as the in-place last stage
    }
    __syncthreads();

    // ---- pair phase: X[k], X[NC-k] from Z; phase advance against the previous frame
    // frames of this batch that exist: gg in [g_lo, g_hi).  Frame -1 (only a = 0, bi = 0, gg = 0) has
    // phi = 0 <=> X = 1, the initial value of "previous".
    const int g_hi = (int)min((long long)G, b - f_first);
    const int g_lo = f_first < 0 ? 1 : 0;
#pragma unroll(kUnrollPair)
    for (int gg = 0; gg < G; ++gg) {
      if (gg >= g_hi) break;
      if (gg < g_lo) continue;
      C* zb = buf + gg * BUF;
      const bool emit = (bi != 0 || gg != 0);  // the chunk's leading halo frame only seeds "previous"
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        const int k = 1 + tid + q * THREADS;
        if (k <= NC / 2) {
          const int mbin = NC - k;
          const C za = zb[fft_pad(k)];
          const C zc = zb[fft_pad(mbin)];
#if MLX_KA_DERIVE_WPAIR
          // pair q's twiddle = pair 0's rotated by exp(-2 pi i q THREADS / N) = q * (16 THREADS / N) sixteenths
          C w = wpair[0];
          constexpr int STEP16 = 16 * THREADS / N;
          static_assert(QP == 1 || (16 * THREADS) % N == 0, "pair twiddles are a whole number of sixteenths apart");
          switch (q * STEP16) {
            case 1: w = rot16_pinned<-1, 1>(w); break;
            case 2: w = rot16_pinned<-1, 2>(w); break;
            case 3: w = rot16_pinned<-1, 3>(w); break;
            case 4: w = rot16_pinned<-1, 4>(w); break;
            case 6: w = rot16_pinned<-1, 6>(w); break;
            default: break;
          }
#else
          const C w = wpair[q];
#endif
          const double er = kSplit * (za.x + zc.x), ei = kSplit * (za.y - zc.y);
          const double dr = kSplit * (za.x - zc.x), di = kSplit * (za.y + zc.y);
          const double tr_ = dr * w.x - di * w.y, ti_ = dr * w.y + di * w.x;
          const C xk{er + ti_, ei - tr_};
          const C xm{er - ti_, -ei - tr_};
          float mag;
          int dq;
          bool flip;
          analysis_bin(xk.x, xk.y, pk[q].x, pk[q].y, ppk[q], pmk[q], k, false, mag, dq, flip);
          const uint32_t rk0 = __float_as_uint(mag) | (flip ? 0x80000000u : 0u), rk1 = (uint32_t)dq;
          analysis_bin(xm.x, xm.y, pm[q].x, pm[q].y, ppm[q], pmm[q], mbin, false, mag, dq, flip);
          // both records of the pair as ONE 16-byte store into slot k (this thread read it above: race-free;
          // 8-byte stores at a 16-byte stride were two-way bank conflicts)
          if (emit)
            *reinterpret_cast<uint4*>(zb + fft_pad(k)) =
                make_uint4(rk0, rk1, __float_as_uint(mag) | (flip ? 0x80000000u : 0u), (uint32_t)dq);
          pk[q] = xk;
          pm[q] = xm;
        }
      }
    }
    if (tid == THREADS - 1) {  // DC and Nyquist are real: X[0] = Re Z0 + Im Z0, X[NC] = Re Z0 - Im Z0
      for (int gg = g_lo; gg < g_hi; ++gg) {
        C* zb = buf + gg * BUF;
        const C z0 = zb[0];
        constexpr double kReal = WD ? 2.0 : 1.0;  // Z is staged at half scale when the window is
        const MagD m0 = analysis_real_bin(kReal * (z0.x + z0.y), pp0, pm0);
        const MagD mn = analysis_real_bin(kReal * (z0.x - z0.y), ppn, pmn);
        if (bi != 0 || gg != 0)  // bins 0 and NC share slot 0
          *reinterpret_cast<uint4*>(zb) = make_uint4(__float_as_uint(m0.mag), (uint32_t)m0.d, __float_as_uint(mn.mag),
                                                     (uint32_t)mn.d);
      }
    }
    __syncthreads();

    // ---- gather phase: bin shift, exact phase increment, chunk-local scan, spill to HBM
#if MLX_GATHER_V2
    if (fast_shift) {
      // rows of frame gg: base pointer of the batch + compile-time offsets; g_cnt = frames before the
      // end of the wave (their running phase is what the next wave starts from)
      const int e_lo = bi == 0 ? 1 : 0;  // the chunk's leading halo frame emits nothing
      const int g_cnt = (int)min((long long)g_hi, wv.we - f_first);
#if MLX_KA_TOTC_DIRECT
      const int g_we = (int)min((long long)G, wv.we - 1 - f_first);  // frame we - 1 inside this batch (or none)
#endif
      uint2* pst = sc.stage + (row0 + (size_t)(f_first - wv.wb)) * NBP + tid;
      const int r_fix = (int)wv.r_fix;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) {
        if (gg >= e_lo && gg < g_hi) {
          const C* zb = buf + gg * BUF;
#pragma unroll
          for (int q = 0; q < QB; ++q) {
            if (tid + q * THREADS < NC) {
              uint32_t inc;
              const float smag = shift_one_bin_v2(zb, scq[q], r_fix, inc);
              lacc[q] += inc;
#if MLX_KA_TOTC_DIRECT
              if (FAST) {
                if (gg == g_we) sc.totc[trow + tid + q * THREADS] = lacc[q];
              } else
#endif
              if (gg == g_cnt - 1) totc[q] = lacc[q];
              pst[gg * NBP + q * THREADS] = make_uint2(__float_as_uint(smag), lacc[q]);
            }
          }
        }
      }
    } else
#endif
#pragma unroll(kUnrollGather)
    for (int gg = (bi == 0 ? 1 : 0); gg < G; ++gg) {
      const long long ff = f_first + gg;
      if (ff >= b) break;
      const C* zb = buf + gg * BUF;
      const size_t row = (row0 + (size_t)(ff - wv.wb)) * NBP;
      float r = wv.rate;
      uint32_t r_fix = (uint32_t)wv.r_fix;  // rate * 2^26 <= 2^28
      if (per_frame_rate) {
        r = tr.rate_pf[ff];
        r_fix = (uint32_t)((double)r * 67108864.0);
      }
      const bool counted = ff < wv.we;
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        const int j = tid + q * THREADS;
        if (j < NC) {
          uint32_t kk, inc;
          if (per_frame_rate) {
            gather_entry_slow(j, r, NC, kk);
          } else {
            kk = gkq[q];
          }
          const float smag = shift_one_bin<NC, BUF>(zb, j, kk, r_fix, inc);
          lacc[q] += inc;
          if (counted) totc[q] = lacc[q];
          sc.stage[row + j] = make_uint2(__float_as_uint(smag), lacc[q]);
        }
      }
    }

    // ---- the Nyquist output bin j = NC does not fit the bins-per-thread tiling: the last warp takes
    //      it for the whole batch, one lane per frame, with a warp scan for the running phase
    if (tid >= THREADS - 32) {
      const int lane = tid & 31;
      for (int g0 = 0; g0 < G; g0 += 32) {
        const int gg = g0 + lane;
        const long long ff = f_first + gg;
        const bool valid = gg < G && gg >= (bi == 0 ? 1 : 0) && ff < b;
        uint32_t inc = 0u;
        float smag = 0.f;
        if (valid) {
          float r = wv.rate;
          uint32_t r_fix = (uint32_t)wv.r_fix, kk;
          if (FAST) {
            smag = shift_one_bin_v2(buf + gg * BUF, make_shift_const_a(NC, gk_nyq, r_fix, NC, true, 16u * ZSLOT),
                                    (int)r_fix, inc);
          } else {
            if (per_frame_rate) {
              r = tr.rate_pf[ff];
              r_fix = (uint32_t)((double)r * 67108864.0);
              gather_entry_slow(NC, r, NC, kk);
            } else {
              kk = gk_nyq;
            }
            smag = shift_one_bin<NC, BUF>(buf + gg * BUF, NC, kk, r_fix, inc);
          }
        }
        uint32_t run = inc;  // inclusive scan over the lanes (= frames, ascending); lanes >= G hold 0
#pragma unroll
        for (int o = 1; o < (G < 32 ? G : 32); o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, run, o);
          if (lane >= o) run += v;
        }
        const uint32_t mine = lacc_nyq + run;
        if (valid) {
          const size_t row = (row0 + (size_t)(ff - wv.wb)) * NBP;
          sc.stage[row + NC] = make_uint2(__float_as_uint(smag), mine);
        }
        // phase at the last frame before the wave end (carried into the next wave)
        const unsigned cm = __ballot_sync(0xffffffffu, valid && ff < wv.we);
        if (cm) totc_nyq = __shfl_sync(0xffffffffu, mine, 31 - __clz(cm));
        lacc_nyq += __shfl_sync(0xffffffffu, run, (G < 32 ? G : 32) - 1);  // the batch total (lanes >= G hold 0)
      }
    }

    // ---- peak bin (lowest k on exact ties) and f0, one warp per frame
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int gg = warp + (bi == 0 ? 1 : 0); gg < G; gg += THREADS / 32) {
        const long long ff = f_first + gg;
        if (ff >= b || ff >= wv.we) break;
        const C* zb = buf + gg * BUF;
        // Non-negative floats order like their bit patterns, so the maximum over the warp is ONE integer
        // warp reduction and "lowest k on exact ties" a second one (10 shuffles + compares before: the four
        // warps doing this keep the other four waiting at the barrier below).  fmaxf(., 0) maps a NaN
        // magnitude to 0, which is what `v > best` did with it.
        uint32_t bb = 0u;
        int bk = wv.kmin;
#pragma unroll 4
        for (int k = wv.kmin + lane; k <= wv.kmax; k += 32) {
          const uint32_t v = __float_as_uint(fmaxf(fabsf(rec(zb, k).mag), 0.f));
          if (v > bb) { bb = v; bk = k; }
        }
        const uint32_t best = __reduce_max_sync(0xffffffffu, bb);
        bk = (int)__reduce_min_sync(0xffffffffu, bb == best ? (uint32_t)bk : 0x7fffffffu);
        if (lane == 0) {
          if (tr.peak) tr.peak[ff] = bk;
          if (tr.f0) {
            const MagD mb = rec(zb, bk);
            long long dd = (long long)mb.d;
            if (__float_as_uint(mb.mag) >> 31) dd += (dd < 0) ? 4294967296LL : -4294967296LL;
            tr.f0[ff] = ((float)bk + (float)dd * 9.313225746154785e-10f) * wv.fs_over_N;  // nu = k + 4 d
          }
        }
      }
    }
    __syncthreads();  // shared buffers are reused by the next batch
  }

#pragma unroll
  for (int q = 0; q < QB; ++q) {
    const int j = tid + q * THREADS;
    if (j < NC) {
      sc.tot[trow + j] = lacc[q];    // all frames of the chunk: prefix of the later chunks of this wave
#if MLX_KA_TOTC_DIRECT
      if (FAST) {  // every frame of the chunk before the wave end (b <= we), none (a >= we), or stored at frame we - 1
        if (b <= wv.we) sc.totc[trow + j] = lacc[q];
        else if (a >= wv.we) sc.totc[trow + j] = 0u;
      } else
#endif
      sc.totc[trow + j] = totc[q];   // frames < we only: what the next wave starts from
    }
  }
  if (tid == THREADS - 1) {
    sc.tot[trow + NC] = lacc_nyq;
    sc.totc[trow + NC] = totc_nyq;
  }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan over the chunk totals of one wave.  Block = 32 bins x 8 chunk segments; every
// thread first reduces its segment (independent loads), the segment sums are combined through
// shared memory, then the segment is rescanned to emit the prefixes.
constexpr int kScanSeg = 8;
__global__ void __launch_bounds__(32 * kScanSeg)
pv_scan_kernel(int nb, int nbp, int nchunks, const PvScratch sc) {
  __shared__ uint32_t s_all[kScanSeg][32], s_cnt[kScanSeg][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const size_t tr = blockIdx.y;
  const int len = (nchunks + kScanSeg - 1) / kScanSeg;
  const int c0 = min(seg * len, nchunks), c1 = min(c0 + len, nchunks);
  const uint32_t* __restrict__ tot = sc.tot + tr * nchunks * (size_t)nbp;
  const uint32_t* __restrict__ totc = sc.totc + tr * nchunks * (size_t)nbp;
  uint32_t* __restrict__ pre = sc.pre + tr * nchunks * (size_t)nbp;
  uint32_t sum = 0u, sumc = 0u;
  if (j < nb) {
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
      sum += __ldg(tot + (size_t)c * nbp + j);
      sumc += __ldg(totc + (size_t)c * nbp + j);
    }
  }
  s_all[seg][lane] = sum;
  s_cnt[seg][lane] = sumc;
  __syncthreads();
  if (j >= nb) return;
  const uint32_t carry = sc.carry[tr * nbp + j];
  uint32_t run = carry;
  for (int q = 0; q < seg; ++q) run += s_all[q][lane];
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    const uint32_t t = __ldg(tot + (size_t)c * nbp + j);
    pre[(size_t)c * nbp + j] = run;
    run += t;
  }
  if (seg == 0) {
    uint32_t counted = carry;
    for (int q = 0; q < kScanSeg; ++q) counted += s_cnt[q][lane];
    sc.carry[tr * nbp + j] = counted;  // phase at the start of the next wave
  }
}

// ------------------------------------------------------------------------------------------------
// K_S
// the reference's export conversion (app.cpp:1209-1212): int16(x * 32767.), double product, truncation
__device__ __forceinline__ short pcm16(float v) { return (short)__double2int_rz((double)v * 32767.); }

// Three synthesis CTAs per SM up to MLX_KS_MINB3_MAXN (80 registers), two beyond.
template <int N>
struct KsTune {
  static constexpr int MINB = N <= MLX_KS_MINB3_MAXN ? 3 : 2;
};

// How the kernel got here (each step measured, profiles/README.md):
//  * the stage records of a batch (consecutive rows = one contiguous span) arrive by ONE TMA bulk copy issued while
//    the previous batch is still in its inverse FFT / overlap-add (the global loads were 17 % of the stall samples);
//  * the synthesis spectrum of a frame is folded by the threads that transform it: thread t builds its own slots
//    k = t + m*TPF (m < 8) in registers and hands the mirrored bins NC - k -- slot 15 - m of thread TPF - t -- over
//    through an unpadded staging area; phase prefixes come through L1 and twiddles by constant rotation (sixteen of
//    each do not fit the 80 registers of three CTAs per SM).  Before, the pair threads wrote all of Z to shared memory
//    and the transform loaded it back: 162 wavefronts per frame against 65 + 32;
//  * the synthesis window is applied by the overlap-add threads: a thread owns the same output columns for every
//    frame, so its 8 window factors per column live in registers and the product fuses into the sum (FFMA).
template <int N, int G, bool O16>
__global__ void __launch_bounds__(PvCfg<N, G>::THREADS, KsTune<N>::MINB)
pv_synth_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = PvCfg<N, G>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, H = Cfg::H, NBP = Cfg::NBP;
  constexpr int THREADS = Cfg::THREADS, BUF = Cfg::BUF;
  constexpr int H2 = H / 2;
  constexpr int COLS = (H2 + THREADS - 1) / THREADS;  // overlap-add columns (float2) per thread
  using C = cplx<float>;
  using F = Fft<float, NC, +1>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                                  // [G][BUF]
  uint2* s_rec = reinterpret_cast<uint2*>(smem_raw + sizeof(C) * G * BUF);  // [G][NBP] stage records of the batch
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_rec + G * NBP);

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  const PvTrack tr = tracks[blockIdx.y];
  if (O16 ? (tr.out16 == nullptr) : (tr.out == nullptr)) return;
  const long long hop_lim = min(wv.we, tr.F);
  const long long a = wv.wb + (long long)blockIdx.x * wv.CS;
  if (a >= hop_lim) return;
  const long long b = min(a + (long long)wv.CS, hop_lim);
  const long long flim = min(b + 3, tr.F);  // frames [a, flim) contribute to hops [a, b)

  FftTwiddles<float, NC, +1, MLX_KS_PRE1 != 0> twd;
  twd.init(t, tb.tw_f);
  const GroupBar<TPF> bar = make_group_bar<TPF>(g, tid);
  const float2 wf0v = __ldg(reinterpret_cast<const float2*>(tb.twr_f + t));  // exp(-2 pi i t / N)
  const C wf0{wf0v.x, wf0v.y};

  const int nbatch = (int)((flim - a + G - 1) / G);
  const int nfr_total = (int)(flim - a);
  const int nhop = (int)(b - a);
  const size_t row0 = (size_t)blockIdx.y * wv.rows + (size_t)(a - wv.wb);
  const int a_off = (int)(a - wv.wb);
  float2 p0[COLS], p1[COLS], p2[COLS];  // pending overlap-add sums of the three youngest hops
  float2 wq[COLS][4];  // synthesis window (gain and 1/N included) at this thread's columns, quarter q of the frame
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    p0[c] = p1[c] = p2[c] = make_float2(0.f, 0.f);
    const int i2 = tid + c * THREADS;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      wq[c][q] = i2 < H2 ? __ldg(reinterpret_cast<const float2*>(tb.wsyn + q * H + 2 * i2)) : make_float2(0.f, 0.f);
  }
  // hop `hrel` (relative to a) of column i2: written only if this chunk owns it
  auto emit_hop = [&](int hrel, int i2, float2 v) {
    if (hrel >= 0 && hrel < nhop) {
      const long long o = (a + hrel) * H + 2 * i2;
      if constexpr (O16) {
        if (o + 1 < tr.n) {
          *reinterpret_cast<short2*>(tr.out16 + o) = make_short2(pcm16(v.x), pcm16(v.y));
        } else if (o < tr.n) {
          tr.out16[o] = pcm16(v.x);
        }
      } else {
        if (o + 1 < tr.n) {
          *reinterpret_cast<float2*>(tr.out + o) = v;
        } else if (o < tr.n) {
          tr.out[o] = v.x;
        }
      }
    }
  };

  // chunks that end inside the track (all but the last one of a track) need no per-sample bounds test
  const bool interior = b * H <= tr.n;
  float* const pout = O16 ? nullptr : tr.out + a * H + 2 * tid;  // column 0 of this thread in hop a
  short* const pout16 = O16 ? tr.out16 + a * H + 2 * tid : nullptr;
  const uint2* const pst = sc.stage + row0 * NBP;
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)min(G, nfr_total) * NBP * (uint32_t)sizeof(uint2);
    mbar_expect_tx(mbar, bytes);
    tma_load_1d(s_rec, pst, bytes, mbar);
  }
  const uint32_t* const pre_trk = sc.pre + (size_t)blockIdx.y * wv.nchunksA * NBP;
  int ca_b = a_off / wv.CA, rem_b = a_off % wv.CA;  // analysis chunk of frame a, and a's position inside it

  for (int bi = 0; bi < nbatch; ++bi) {
    const int fb = bi * G;                       // frame index relative to a
    const int nfr = min(G, nfr_total - fb);      // frames present in this batch
    mbar_wait(mbar, bi & 1);
    float* const pob = pout + (long long)(fb - 3) * H;  // hop fb - 3: where frame fb's first quarter completes
    short* const pob16 = pout16 + (long long)(fb - 3) * H;

    // ---- synthesis spectrum Y = smag e^{i theta} of frame fb + g, folded for the N/2-point complex inverse by
    //      the group that transforms it.  Pair (k, NC - k), k = t + m*TPF: Z[k] is slot m of this thread,
    //      Z[NC - k] slot 15 - m of thread TPF - t (slot 16 - m of thread 0 for t = 0): staged at
    //      (8 - m)*TPF - t, where its owner reads (slot - 8)*TPF + t.  Thread 0's pair m = 0 is DC / Nyquist (one
    //      real pair -> Z[0]); it also folds the self-paired bin NC/2 (its slot 8, staged at 0).
    C x[16];
    if (g < nfr) {
      // analysis chunk of this frame -> its phase prefix row (L1-resident).  The chunk index of the batch's first
      // frame is carried along (ca_b, rem_b): no division per frame
      int ca = ca_b;
      for (int r = rem_b + g; r >= wv.CA; r -= wv.CA) ++ca;  // (at most one step unless the chunks are shorter than a batch)
      const uint32_t* pp = pre_trk + (size_t)ca * NBP;
      const uint2* src = s_rec + g * NBP;  // one 8-byte record per bin: .x = shifted magnitude, .y = chunk-local phase sum
      C* zb = buf + g * BUF;
      constexpr SpecRot32 rot = spec_rot32();
      auto fold = [&](uint2 rk, uint2 rm, uint32_t pk, uint32_t pm, C w, C& zk, C& zm) {
        float sk, ck, sm, cm;
        sincos_turns(pk + rk.y, sk, ck);
        sincos_turns(pm + rm.y, sm, cm);
        const float mkq = __uint_as_float(rk.x), mmq = __uint_as_float(rm.x);
        const float ykr = mkq * ck, yki = mkq * sk, ymr = mmq * cm, ymi = mmq * sm;
        // A = Y_k, B = conj(Y_m): E2 = A + B, D2 = A - B, O2 = D2 * conj(W^k)
        const float er = ykr + ymr, ei = yki - ymi;
        const float dr = ykr - ymr, di = yki + ymi;
        const float orr = dr * w.x + di * w.y, oi = di * w.x - dr * w.y;
        zk = C{er - oi, ei + orr};
        zm = C{er + oi, orr - ei};
      };
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int k = t + m * TPF, mbin = NC - k;
        const uint2 rk = src[k], rm = src[mbin];
        const uint32_t pk = __ldg(pp + k), pm = __ldg(pp + mbin);
        // exp(-2 pi i k / N) = w0 * exp(-2 pi i m / 32): one rotation per pair instead of eight registers (the
        // products are pinned inside the loop: hoisted they would be fourteen registers again)
        C w = wf0;
        if (m != 0)
          w = C{fmaf(-wf0.y, rot.s[m], spec_mul_here(wf0.x, rot.c[m])), fmaf(wf0.y, rot.c[m], spec_mul_here(wf0.x, rot.s[m]))};
        C zk, zm;
        fold(rk, rm, pk, pm, w, zk, zm);
        int idx = (8 - m) * TPF - t;
        if (m == 0 && t == 0) {
          // DC / Nyquist: Im forced to 0 (the fold above ran on the pair (0, NC) and is discarded)
          float s0, c0, sn, cn;
          sincos_turns(pk + rk.y, s0, c0);
          sincos_turns(pm + rm.y, sn, cn);
          const float y0 = __uint_as_float(rk.x) * c0, yn = __uint_as_float(rm.x) * cn;
          zk = C{y0 + yn, y0 - yn};
          const uint2 rh = src[NC / 2];
          const uint32_t ph = __ldg(pp + NC / 2);
          const float2 wh = __ldg(reinterpret_cast<const float2*>(tb.twr_f + NC / 2));
          C zdummy;
          fold(rh, rh, ph, ph, C{wh.x, wh.y}, zm, zdummy);  // bin NC/2 pairs with itself
          idx = 0;
        }
        x[m] = zk;
        zb[idx] = zm;
      }
    }
    __syncthreads();
    if (tid == 0 && bi + 1 < nbatch) {  // every thread has taken its records out of s_rec: refill for batch bi + 1
      const uint32_t bytes = (uint32_t)min(G, nfr_total - (fb + G)) * NBP * (uint32_t)sizeof(uint2);
      fence_proxy_async();
      mbar_expect_tx(mbar, bytes);
      tma_load_1d(s_rec, pst + (size_t)(fb + G) * NBP, bytes, mbar);
    }

    // ---- inverse FFT, result in place as real pairs (the synthesis window is applied by the overlap-add)
    if (g < nfr) {
      C* zb = buf + g * BUF;
#pragma unroll
      for (int m = 8; m < 16; ++m) x[m] = zb[(m - 8) * TPF + t];  // the mirrored bins, staged by the partner thread
      bar.sync();  // every thread of the group holds its inputs: the first stage may overwrite the buffer
      F::run(x, zb, t, twd, bar);
      F::store(x, zb, t);
    }
    __syncthreads();

    // ---- overlap-add, atomics-free and in a fixed order.  Hop h = samples [hH, (h+1)H) is
    //      ((y_h[3] + y_{h+1}[2]) + y_{h+2}[1]) + y_{h+3}[0]  (y_f[q] = windowed quarter q of frame f).
    //      Every thread owns output columns (float2 at 2*i2 inside the hop) for the whole chunk and
    //      keeps the three pending partial sums in registers; frames arrive in ascending order:
    //        emit hop f-3 = p0 + y_f[0];  p0 = p1 + y_f[1];  p1 = p2 + y_f[2];  p2 = y_f[3]
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int i2 = tid + c * THREADS;
      if (i2 < H2) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
          if (gi < nfr) {
            const C* yb = buf + gi * BUF;
            const C q0 = yb[fft_pad(i2)], q1 = yb[fft_pad(H2 + i2)];
            const C q2 = yb[fft_pad(2 * H2 + i2)], q3 = yb[fft_pad(3 * H2 + i2)];
            const float2 o = make_float2(fmaf(q0.x, wq[c][0].x, p0[c].x), fmaf(q0.y, wq[c][0].y, p0[c].y));
            p0[c] = make_float2(fmaf(q1.x, wq[c][1].x, p1[c].x), fmaf(q1.y, wq[c][1].y, p1[c].y));
            p1[c] = make_float2(fmaf(q2.x, wq[c][2].x, p2[c].x), fmaf(q2.y, wq[c][2].y, p2[c].y));
            p2[c] = make_float2(q3.x * wq[c][3].x, q3.y * wq[c][3].y);
            if (interior) {
              const int hrel = fb + gi - 3;
              if (hrel >= 0 && hrel < nhop) {
                if constexpr (O16)
                  *reinterpret_cast<short2*>(pob16 + gi * H + 2 * c * THREADS) = make_short2(pcm16(o.x), pcm16(o.y));
                else
                  *reinterpret_cast<float2*>(pob + gi * H + 2 * c * THREADS) = o;
              }
            } else {
              emit_hop(fb + gi - 3, i2, o);
            }
          }
        }
      }
    }
    __syncthreads();
    rem_b += G;
    while (rem_b >= wv.CA) {
      rem_b -= wv.CA;
      ++ca_b;
    }
  }
  // hops whose later frames do not exist (end of the track): what has been summed is the result
  const int last = nfr_total - 1;
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int i2 = tid + c * THREADS;
    if (i2 < H2) {
      emit_hop(last - 2, i2, p0[c]);
      emit_hop(last - 1, i2, p1[c]);
      emit_hop(last, i2, p2[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K_S32: the synthesis kernel with THIRTY-TWO points per thread, for fftN = 2048 (NC = 1024 = 32 x 32).
// One warp transforms one frame in two radix-32 stages with ONE exchange through shared memory (the 16-point plan
// needs two: 16 x 16 x 4), synchronised by __syncwarp alone; the 31 powers of the second stage's twiddle stay in
// registers for the whole kernel (exact table values, no power tree per frame).  CTA = 4 warps = 4 frames, three
// CTAs per SM at up to 168 registers.  Fold, record staging by TMA, overlap-add and output are those of
// pv_synth_kernel with TPF = 32 (pairs k = t + 32 m, m < 16; mirrored bins staged at (16 - m)*32 - t).
#ifndef MLX_KS32
#define MLX_KS32 1
#endif
struct Rot64 {  // exp(-2 pi i m / 64), m < 16  (c = cos, s = -sin)
  float c[16], s[16];
};
MLX_HDC Rot64 make_rot64() {
  Rot64 r{};
  for (int m = 0; m < 16; ++m) {
    r.c[m] = (float)fft_ccos(6.283185307179586476925286766559 * m / 64.0);
    r.s[m] = (float)-fft_csin(6.283185307179586476925286766559 * m / 64.0);
  }
  return r;
}

template <int N, bool O16>
__global__ void __launch_bounds__(128, 3)
pv_synth32_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  constexpr int NC = N / 2, TPF = 32, G = 4, THREADS = 128, H = N / 4, H2 = H / 2, NBP = NC + 32;
  constexpr int BUF = Fft32x32::BUF;
  constexpr int COLS = H2 / THREADS;
  static_assert(NC == 1024, "two radix-32 stages");
  using C = cplx<float>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);  // [G][BUF]
  uint2* s_rec = reinterpret_cast<uint2*>(buf + G * BUF);  // [G][NBP]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_rec + G * NBP);

  const int tid = threadIdx.x;
  const int g = tid >> 5, t = tid & 31;
  const PvTrack tr = tracks[blockIdx.y];
  if (O16 ? (tr.out16 == nullptr) : (tr.out == nullptr)) return;
  const long long hop_lim = min(wv.we, tr.F);
  const long long a = wv.wb + (long long)blockIdx.x * wv.CS;
  if (a >= hop_lim) return;
  const long long b = min(a + (long long)wv.CS, hop_lim);
  const long long flim = min(b + 3, tr.F);  // frames [a, flim) contribute to hops [a, b)

  // powers of the second stage's twiddle exp(+2 pi i t / NC): exact table values, kept for the whole kernel
  C p1[31];
#pragma unroll
  for (int r = 1; r < 32; ++r) {
    const float2 q = __ldg(reinterpret_cast<const float2*>(tb.tw_f + ((t * r) & (NC - 1))));  // exp(-2 pi i t r / NC)
    p1[r - 1] = C{q.x, -q.y};
  }
  const float2 wf0v = __ldg(reinterpret_cast<const float2*>(tb.twr_f + t));  // exp(-2 pi i t / N)
  const C wf0{wf0v.x, wf0v.y};

  const int nbatch = (int)((flim - a + G - 1) / G);
  const int nfr_total = (int)(flim - a);
  const int nhop = (int)(b - a);
  const size_t row0 = (size_t)blockIdx.y * wv.rows + (size_t)(a - wv.wb);
  const int a_off = (int)(a - wv.wb);
  float2 p0[COLS], pa[COLS], pb[COLS];  // pending overlap-add sums of the three youngest hops
  float2 wq[COLS][4];                   // synthesis window at this thread's columns, quarter q of the frame
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    p0[c] = pa[c] = pb[c] = make_float2(0.f, 0.f);
    const int i2 = tid + c * THREADS;
#pragma unroll
    for (int q = 0; q < 4; ++q) wq[c][q] = __ldg(reinterpret_cast<const float2*>(tb.wsyn + q * H + 2 * i2));
  }
  auto emit_hop = [&](int hrel, int i2, float2 v) {  // hop `hrel` (relative to a) of column i2, if this chunk owns it
    if (hrel >= 0 && hrel < nhop) {
      const long long o = (a + hrel) * H + 2 * i2;
      if constexpr (O16) {
        if (o + 1 < tr.n) *reinterpret_cast<short2*>(tr.out16 + o) = make_short2(pcm16(v.x), pcm16(v.y));
        else if (o < tr.n) tr.out16[o] = pcm16(v.x);
      } else {
        if (o + 1 < tr.n) *reinterpret_cast<float2*>(tr.out + o) = v;
        else if (o < tr.n) tr.out[o] = v.x;
      }
    }
  };
  const bool interior = b * H <= tr.n;
  float* const pout = O16 ? nullptr : tr.out + a * H + 2 * tid;
  short* const pout16 = O16 ? tr.out16 + a * H + 2 * tid : nullptr;
  const uint2* const pst = sc.stage + row0 * NBP;
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)min(G, nfr_total) * NBP * (uint32_t)sizeof(uint2);
    mbar_expect_tx(mbar, bytes);
    tma_load_1d(s_rec, pst, bytes, mbar);
  }
  const uint32_t* const pre_trk = sc.pre + (size_t)blockIdx.y * wv.nchunksA * NBP;
  int ca_b = a_off / wv.CA, rem_b = a_off % wv.CA;  // analysis chunk of frame a, and a's position inside it

  for (int bi = 0; bi < nbatch; ++bi) {
    const int fb = bi * G;
    const int nfr = min(G, nfr_total - fb);
    mbar_wait(mbar, bi & 1);
    float* const pob = pout + (long long)(fb - 3) * H;
    short* const pob16 = pout16 + (long long)(fb - 3) * H;
    C x[32];
    C* const zb = buf + g * BUF;
    if (g < nfr) {
      int ca = ca_b;
      for (int r = rem_b + g; r >= wv.CA; r -= wv.CA) ++ca;
      const uint32_t* pp = pre_trk + (size_t)ca * NBP;
      const uint2* src = s_rec + g * NBP;
      constexpr Rot64 rot = make_rot64();
      auto fold = [&](uint2 rk, uint2 rm, uint32_t pk, uint32_t pm, C w, C& zk, C& zm) {
        float sk, ck, sm, cm;
        sincos_turns(pk + rk.y, sk, ck);
        sincos_turns(pm + rm.y, sm, cm);
        const float mkq = __uint_as_float(rk.x), mmq = __uint_as_float(rm.x);
        const float ykr = mkq * ck, yki = mkq * sk, ymr = mmq * cm, ymi = mmq * sm;
        const float er = ykr + ymr, ei = yki - ymi;
        const float dr = ykr - ymr, di = yki + ymi;
        const float orr = dr * w.x + di * w.y, oi = di * w.x - dr * w.y;
        zk = C{er - oi, ei + orr};
        zm = C{er + oi, orr - ei};
      };
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int k = t + m * TPF, mbin = NC - k;
        const uint2 rk = src[k], rm = src[mbin];
        const uint32_t pk = __ldg(pp + k), pm = __ldg(pp + mbin);
        C w = wf0;  // exp(-2 pi i k / N) = w0 * exp(-2 pi i m / 64)
        if (m != 0)
          w = C{fmaf(-wf0.y, rot.s[m], spec_mul_here(wf0.x, rot.c[m])), fmaf(wf0.y, rot.c[m], spec_mul_here(wf0.x, rot.s[m]))};
        C zk, zm;
        fold(rk, rm, pk, pm, w, zk, zm);
        int idx = (16 - m) * TPF - t;
        if (m == 0 && t == 0) {  // DC / Nyquist (Im forced to 0) and the self-paired bin NC/2 (slot 16 of thread 0)
          float s0, c0, sn, cn;
          sincos_turns(pk + rk.y, s0, c0);
          sincos_turns(pm + rm.y, sn, cn);
          const float y0 = __uint_as_float(rk.x) * c0, yn = __uint_as_float(rm.x) * cn;
          zk = C{y0 + yn, y0 - yn};
          const uint2 rh = src[NC / 2];
          const uint32_t ph = __ldg(pp + NC / 2);
          const float2 wh = __ldg(reinterpret_cast<const float2*>(tb.twr_f + NC / 2));
          C zdummy;
          fold(rh, rh, ph, ph, C{wh.x, wh.y}, zm, zdummy);
          idx = 0;
        }
        x[m] = zk;
        zb[idx] = zm;
      }
    }
    __syncthreads();  // records consumed, staged halves visible
    if (tid == 0 && bi + 1 < nbatch) {
      const uint32_t bytes = (uint32_t)min(G, nfr_total - (fb + G)) * NBP * (uint32_t)sizeof(uint2);
      fence_proxy_async();
      mbar_expect_tx(mbar, bytes);
      tma_load_1d(s_rec, pst + (size_t)(fb + G) * NBP, bytes, mbar);
    }
    if (g < nfr) {
#pragma unroll
      for (int m = 16; m < 32; ++m) x[m] = zb[(m - 16) * TPF + t];
      __syncwarp();  // the staging area is overwritten by the first stage's stores
      Fft32x32::stage0<+1>(x, zb, t);
      __syncwarp();
      Fft32x32::stage1<+1>(x, zb, t, p1);
      Fft32x32::store(x, zb, t);  // natural order: the slots this thread has just read
    }
    __syncthreads();

    // overlap-add as in pv_synth_kernel (window applied here, fixed ascending-frame order)
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int i2 = tid + c * THREADS;
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        if (gi < nfr) {
          const C* yb = buf + gi * BUF;
          const C q0 = yb[pad32(i2)], q1 = yb[pad32(H2 + i2)];
          const C q2 = yb[pad32(2 * H2 + i2)], q3 = yb[pad32(3 * H2 + i2)];
          const float2 o = make_float2(fmaf(q0.x, wq[c][0].x, p0[c].x), fmaf(q0.y, wq[c][0].y, p0[c].y));
          p0[c] = make_float2(fmaf(q1.x, wq[c][1].x, pa[c].x), fmaf(q1.y, wq[c][1].y, pa[c].y));
          pa[c] = make_float2(fmaf(q2.x, wq[c][2].x, pb[c].x), fmaf(q2.y, wq[c][2].y, pb[c].y));
          pb[c] = make_float2(q3.x * wq[c][3].x, q3.y * wq[c][3].y);
          if (interior) {
            const int hrel = fb + gi - 3;
            if (hrel >= 0 && hrel < nhop) {
              if constexpr (O16)
                *reinterpret_cast<short2*>(pob16 + gi * H + 2 * c * THREADS) = make_short2(pcm16(o.x), pcm16(o.y));
              else
                *reinterpret_cast<float2*>(pob + gi * H + 2 * c * THREADS) = o;
            }
          } else {
            emit_hop(fb + gi - 3, i2, o);
          }
        }
      }
    }
    __syncthreads();
    rem_b += G;
    while (rem_b >= wv.CA) {
      rem_b -= wv.CA;
      ++ca_b;
    }
  }
  const int last = nfr_total - 1;
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int i2 = tid + c * THREADS;
    emit_hop(last - 2, i2, p0[c]);
    emit_hop(last - 1, i2, pa[c]);
    emit_hop(last, i2, pb[c]);
  }
}
constexpr size_t kSmemS32 = sizeof(cplx<float>) * 4 * (1024 + 32) + sizeof(uint2) * 4 * (1024 + 32) + 64;

// ------------------------------------------------------------------------------------------------
template <int N>
static cudaError_t configure_n() {
  constexpr int G = PvG<N>::value, GA = PvG<N>::analyze;
  cudaError_t e = cudaSuccess;
  // the analysis kernel lives on shared memory: ask for the largest carve-out so that MLX_KA_CTAS
  // CTAs are resident
  for (const void* fn : {(const void*)pv_analyze_kernel<N, GA, false>, (const void*)pv_analyze_kernel<N, GA, true>}) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PvCfg<N, GA>::SMEM_A);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(pv_synth_kernel<N, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)PvCfg<N, G>::SMEM_S);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pv_synth_kernel<N, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)PvCfg<N, G>::SMEM_S);
  if (e != cudaSuccess) return e;
  if constexpr (N == 2048 && MLX_KS32) {
    e = cudaFuncSetAttribute(pv_synth32_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemS32);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pv_synth32_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemS32);
  }
  return e;
}

#define MLX_PV_DISPATCH(N_, ...)                  \
  switch (N_) {                                   \
    case 512: { constexpr int N = 512; __VA_ARGS__; } break;   \
    case 1024: { constexpr int N = 1024; __VA_ARGS__; } break; \
    case 2048: { constexpr int N = 2048; __VA_ARGS__; } break; \
    case 4096: { constexpr int N = 4096; __VA_ARGS__; } break; \
    case 8192: { constexpr int N = 8192; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;        \
  }

cudaError_t pv_configure(int fftN) {
  MLX_PV_DISPATCH(fftN, return configure_n<N>());
  return cudaSuccess;
}
int pv_group_count(int fftN) { return 8192 / fftN; }
int pv_group_count_analyze(int fftN) {
  switch (fftN) {
    case 512: return PvG<512>::analyze;
    case 1024: return PvG<1024>::analyze;
    case 2048: return PvG<2048>::analyze;
    case 4096: return PvG<4096>::analyze;
    case 8192: return PvG<8192>::analyze;
  }
  return 1;
}
int pv_threads(int) { return 256; }
size_t pv_analyze_smem(int fftN) {
  switch (fftN) {
    case 512: return PvCfg<512, PvG<512>::analyze>::SMEM_A;
    case 1024: return PvCfg<1024, PvG<1024>::analyze>::SMEM_A;
    case 2048: return PvCfg<2048, PvG<2048>::analyze>::SMEM_A;
    case 4096: return PvCfg<4096, PvG<4096>::analyze>::SMEM_A;
    case 8192: return PvCfg<8192, PvG<8192>::analyze>::SMEM_A;
  }
  return 0;
}
size_t pv_synth_smem(int fftN) {
  switch (fftN) {
    case 512: return PvCfg<512, 16>::SMEM_S;
    case 1024: return PvCfg<1024, 8>::SMEM_S;
    case 2048: return PvCfg<2048, 4>::SMEM_S;
    case 4096: return PvCfg<4096, 2>::SMEM_S;
    case 8192: return PvCfg<8192, 1>::SMEM_S;
  }
  return 0;
}

cudaError_t launch_pv_analyze(int fftN, const PvTrack* tracks, int ntracks, const PvWave& wv,
                              const PvTables& tb, const PvScratch& sc, bool constant_rate_up, cudaStream_t st) {
  MLX_PV_DISPATCH(fftN, {
    constexpr int G = PvG<N>::analyze;
    dim3 grid(wv.nchunksA, ntracks);
    if (constant_rate_up)
      pv_analyze_kernel<N, G, true><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_A, st>>>(tracks, wv, tb, sc);
    else
      pv_analyze_kernel<N, G, false><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_A, st>>>(tracks, wv, tb, sc);
  });
  return cudaGetLastError();
}

cudaError_t launch_pv_scan(int fftN, int ntracks, const PvWave& wv, const PvScratch& sc, cudaStream_t st) {
  const int nb = fftN / 2 + 1;
  dim3 grid((nb + 31) / 32, ntracks);
  pv_scan_kernel<<<grid, 32 * kScanSeg, 0, st>>>(nb, pv_nbp(fftN), wv.nchunksA, sc);
  return cudaGetLastError();
}

cudaError_t launch_pv_synth(int fftN, const PvTrack* tracks, int ntracks, const PvWave& wv,
                            const PvTables& tb, const PvScratch& sc, bool out16, cudaStream_t st) {
#if MLX_KS32
  static const bool ks32 = getenv("MLX_PV_NO_KS32") == nullptr;  // (A/B switch; the 32-point plan is the default at 2048)
  if (fftN == 2048 && ks32) {
    dim3 grid(wv.nchunksS, ntracks);
    if (out16) pv_synth32_kernel<2048, true><<<grid, 128, kSmemS32, st>>>(tracks, wv, tb, sc);
    else pv_synth32_kernel<2048, false><<<grid, 128, kSmemS32, st>>>(tracks, wv, tb, sc);
    return cudaGetLastError();
  }
#endif
  MLX_PV_DISPATCH(fftN, {
    constexpr int G = PvG<N>::value;
    dim3 grid(wv.nchunksS, ntracks);
    if (out16)
      pv_synth_kernel<N, G, true><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_S, st>>>(tracks, wv, tb, sc);
    else
      pv_synth_kernel<N, G, false><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_S, st>>>(tracks, wv, tb, sc);
  });
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// inspection of a staged analysis (mlx_pv_stage_export_dev): what K_S will synthesise from
__global__ void __launch_bounds__(256) pv_stage_export_kernel(int nb, int nbp, int track, const PvWave wv,
                                                              const PvScratch sc, long long f0, long long count,
                                                              float* __restrict__ smag, uint32_t* __restrict__ phase) {
  const long long fr = blockIdx.y;  // frame f0 + fr
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (fr >= count || j >= nb) return;
  const long long rel = f0 + fr - wv.wb;  // row of the wave
  const int chunk = (int)(rel / wv.CA);
  const uint2 rec = sc.stage[((size_t)track * wv.rows + (size_t)rel) * nbp + j];
  const uint32_t pre = sc.pre[((size_t)track * wv.nchunksA + chunk) * nbp + j];
  smag[fr * nb + j] = __uint_as_float(rec.x);
  phase[fr * nb + j] = pre + rec.y;
}

cudaError_t launch_pv_stage_export(int fftN, int track, const PvWave& wv, const PvScratch& sc, long long f0,
                                   long long count, float* smag, uint32_t* phase, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  const int nb = fftN / 2 + 1;
  for (long long done = 0; done < count; done += 32768) {  // gridDim.y limit
    const long long c = count - done < 32768 ? count - done : 32768;
    dim3 grid((nb + 255) / 256, (unsigned)c);
    pv_stage_export_kernel<<<grid, 256, 0, st>>>(nb, pv_nbp(fftN), track, wv, sc, f0 + done, c, smag + done * nb,
                                                 phase + done * nb);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// int16 PCM -> float, 8 samples per thread (one 16-byte load, two 16-byte stores); x = s * 2^-15 exactly
__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const short* __restrict__ in, float* __restrict__ out,
                                                             long long n) {
  const long long i8 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i8 + 8 <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(in + i8));
    const int w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f[2 * q] = (float)(short)(w[q] & 0xffff) * 3.0517578125e-05f;
      f[2 * q + 1] = (float)(w[q] >> 16) * 3.0517578125e-05f;
    }
    *reinterpret_cast<float4*>(out + i8) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(out + i8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    for (long long i = i8; i < min(i8 + 8, n); ++i) out[i] = (float)in[i] * 3.0517578125e-05f;
  }
}

cudaError_t launch_pcm16_to_float(const short* in, float* out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const long long threads = (n + 7) / 8;
  pcm16_to_float_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace mlx
