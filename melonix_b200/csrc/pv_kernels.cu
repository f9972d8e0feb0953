// melonix_b200/csrc/pv_kernels.cu -- phase-vocoder kernels for sm_100a (PV-spec v1).
//
// NOT IN REFERENCE: melonix has no phase vocoder; these kernels implement the
// north-star's analysis / pitch-detect / bin-shift / resynthesis path.  Frame geometry follows the
// reference's Spec jobs (spec.cpp:47, spec-cache.cpp:63-65).
//
//   pv_analyze_kernel  K_A  one CTA = one chunk of consecutive frames of one track, G frames per
//                           batch.  TMA bulk load of the batch's sample tile -> Hann window ->
//                           FP64 real FFT (N/2-point complex Stockham in shared memory) ->
//                           d = arg(X_f conj(X_{f-1}) (-i)^k) -> magnitude, peak bin, f0 ->
//                           bin-shift gather -> exact uint32 phase increments, chunk-local scan.
//   pv_scan_kernel          exclusive scan of the chunk totals per (track, bin).
//   pv_synth_kernel    K_S  theta = prefix + local sum -> Y = smag e^{i theta} -> FP32 inverse real
//                           FFT -> synthesis window -> atomics-free overlap-add in shared memory
//                           (fixed ascending-frame summation order) -> float4 stores.
//
// Why FP64 in K_A: the wrapped phase difference has a cut at +-pi; a frame whose FP32 spectrum
// lands on the other side of the cut than the oracle's shifts that bin's accumulated phase by
// frac(rate) turns for the rest of the track.  B200 runs FP64 at half the FP32 rate, which makes
// the analysis transform robust for ~2x its FP32 cost.  Everything after the cut decision is FP32 /
// exact uint32.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "fft.cuh"
#include "kernels.h"

namespace mlx {

// ------------------------------------------------------------------------------------------------
template <int N, int G>
struct PvCfg {
  static constexpr int NC = N / 2;
  static constexpr int TPF = NC / 16;
  static constexpr int H = N / 4;
  static constexpr int NB = NC + 1;
  static constexpr int NBP = NC + 32;
  static constexpr int THREADS = G * TPF;
  static constexpr int BUF = FftPlan<NC>::BUF;
  static constexpr int TILE = N + (G - 1) * H;  // floats per batch tile
  static constexpr int QP = (NC / 2 + 1 + THREADS - 1) / THREADS;  // pair slots per thread
  static constexpr int QB = (NB + THREADS - 1) / THREADS;          // bin slots per thread
  static constexpr size_t SMEM_A = sizeof(cplx<double>) * G * BUF + sizeof(float) * TILE +
                                   sizeof(float) * 2 * G * NBP + 64;
  static constexpr size_t SMEM_S = sizeof(cplx<float>) * G * BUF + sizeof(float) * 2 * 3 * H + 64;
};

// G such that G * (N/2) = 4096 complex points per batch -> 256 threads for every N.
template <int N>
struct PvG {
  static constexpr int value = 8192 / N;
};

template <int TPF>
struct GroupBar {
  int id;
  unsigned mask;
  __device__ __forceinline__ void sync() const {
    if constexpr (TPF < 32) {
      __syncwarp(mask);
    } else if constexpr (TPF == 32) {
      __syncwarp();
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(TPF) : "memory");
    }
  }
};

// named barrier 1+g for groups of >= 64 threads; lane mask of the group for sub-warp groups
template <int TPF>
__device__ __forceinline__ GroupBar<TPF> make_group_bar(int g, int tid) {
  unsigned mask = 0xffffffffu;
  if constexpr (TPF < 32) mask = ((1u << TPF) - 1u) << (((tid & 31) / TPF) * TPF);
  return GroupBar<TPF>{1 + g, mask};
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// trunc(float(k) * r) with a plain float multiply, exactly as the spec (A.5) and the oracle do.
__device__ __forceinline__ int shift_bin(int k, float r) { return (int)truncf(__fmul_rn((float)k, r)); }

// K_j = { k in [0, NC] : shift_bin(k, r) == j };  returns klo > khi when empty.
__device__ __forceinline__ void gather_range(int j, float r, int NC, int& klo, int& khi) {
  int k = (int)(__fdividef((float)j, r)) - 1;
  k = max(0, min(k, NC));
  while (k <= NC && shift_bin(k, r) < j) ++k;
  while (k > 0 && shift_bin(k - 1, r) >= j) --k;
  klo = k;
  if (k > NC || shift_bin(k, r) != j) {
    khi = k - 1;
    return;
  }
  khi = k;
  while (khi + 1 <= NC && shift_bin(khi + 1, r) == j) ++khi;
}

// ------------------------------------------------------------------------------------------------
// K_A
template <int N, int G>
__global__ void __launch_bounds__(PvCfg<N, G>::THREADS, 1)
pv_analyze_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = PvCfg<N, G>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, H = Cfg::H, NB = Cfg::NB, NBP = Cfg::NBP;
  constexpr int THREADS = Cfg::THREADS, BUF = Cfg::BUF, TILE = Cfg::TILE, QP = Cfg::QP, QB = Cfg::QB;
  using C = cplx<double>;
  using F = Fft<double, NC, -1>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                       // [G][BUF]
  float* tile = reinterpret_cast<float*>(buf + G * BUF);         // [TILE]
  float* s_mag = tile + TILE;                                    // [G][NBP]
  float* s_del = s_mag + G * NBP;                                // [G][NBP]  d / (2 pi), turns
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_del + G * NBP); // 8-byte aligned (all sizes even)

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  const PvTrack tr = tracks[blockIdx.y];
  const long long lim = min(wv.we + 3, tr.F);  // analysis runs three frames past the owned window
  const long long a = wv.wb + (long long)blockIdx.x * wv.CA;
  const size_t trow = ((size_t)blockIdx.y * wv.nchunksA + blockIdx.x) * NBP;
  if (a >= lim) {  // chunk past the end of this track: contributes nothing to the scan
    for (int j = tid; j < NB; j += THREADS) sc.tot[trow + j] = 0u;
    return;
  }
  const long long b = min(a + (long long)wv.CA, lim);

  if (tid == 0) mbar_init(mbar, 1);

  FftTwiddles<double, NC, -1> twd;
  twd.init(t, tb.tw_d);
  float wreg[32];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 w2 = *reinterpret_cast<const float2*>(tb.win + 2 * (t + m * TPF));
    wreg[2 * m] = w2.x;
    wreg[2 * m + 1] = w2.y;
  }
  // pair slots: bins k and NC-k, k = tid + q*THREADS in [0, NC/2]
  C wr[QP], pk[QP], pm[QP];
#pragma unroll
  for (int q = 0; q < QP; ++q) {
    const int k = tid + q * THREADS;
    wr[q] = (k <= NC / 2) ? tb.twr_d[k] : C{1.0, 0.0};
    pk[q] = C{1.0, 0.0};
    pm[q] = C{1.0, 0.0};
  }
  uint32_t lacc[QB], tot[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) lacc[q] = tot[q] = 0u;

  const GroupBar<TPF> bar = make_group_bar<TPF>(g, tid);
  __syncthreads();  // mbarrier initialised

  const int nbatch = (int)((b - a + 1 + G - 1) / G);  // frames a-1 .. b-1
  const size_t row0 = (size_t)blockIdx.y * wv.rows;
  uint32_t parity = 0;

  for (int bi = 0; bi < nbatch; ++bi) {
    const long long f_first = a - 1 + (long long)bi * G;
    // ---- TMA: samples [(f_first-3)H, (f_first+G)H) of the zero-padded track
    if (tid == 0) {
      mbar_expect_tx(mbar, TILE * sizeof(float));
      tma_load_1d(tile, tr.x + (f_first - 3) * H, TILE * sizeof(float), mbar);
    }
    mbar_wait(mbar, parity);
    parity ^= 1u;

    // ---- forward FP64 real FFT of frame f_first + g (N/2-point complex on even/odd samples)
    const long long fg = f_first + g;
    if (fg >= 0 && fg < b) {
      C x[16];
      const float* src = tile + g * H;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const float2 s2 = *reinterpret_cast<const float2*>(src + 2 * (t + m * TPF));
        x[m] = C{(double)wreg[2 * m] * (double)s2.x, (double)wreg[2 * m + 1] * (double)s2.y};
      }
      F::run(x, buf + g * BUF, t, twd, bar);
      F::store(x, buf + g * BUF, t);  // same thread-private slots as the in-place last stage
    }
    __syncthreads();

    // ---- pair phase: X[k], X[NC-k] from Z; phase advance against the previous frame
#pragma unroll 1
    for (int gg = 0; gg < G; ++gg) {
      const long long ff = f_first + gg;
      if (ff >= b) break;
      if (ff < 0) continue;  // frame -1: phi = 0 <=> X = 1 (initial pk/pm)
      const C* zb = buf + gg * BUF;
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        const int k = tid + q * THREADS;
        if (k > NC / 2) continue;
        const int mbin = NC - k;
        const C za = zb[fft_pad(k)];
        const C zc = zb[fft_pad(mbin & (NC - 1))];
        C xk, xm;
        if (k == 0) {
          xk = C{za.x + za.y, 0.0};
          xm = C{za.x - za.y, 0.0};
        } else {
          const double er = 0.5 * (za.x + zc.x), ei = 0.5 * (za.y - zc.y);
          const double dr = 0.5 * (za.x - zc.x), di = 0.5 * (za.y + zc.y);
          const double tr_ = dr * wr[q].x - di * wr[q].y, ti_ = dr * wr[q].y + di * wr[q].x;
          xk = C{er + ti_, ei - tr_};
          xm = C{er - ti_, -ei - tr_};
        }
        if (bi != 0 || gg != 0) {  // the chunk's leading halo frame only seeds pk/pm
          // Z = X conj(Xprev) (-i)^bin ; real bins (0 and NC) have Im Z := +0
          {
            double zr = xk.x * pk[q].x + xk.y * pk[q].y, zi = xk.y * pk[q].x - xk.x * pk[q].y;
            double rr, ri;
            switch (k & 3) {
              case 1: rr = zi; ri = -zr; break;
              case 2: rr = -zr; ri = -zi; break;
              case 3: rr = -zi; ri = zr; break;
              default: rr = zr; ri = zi; break;
            }
            if (k == 0) ri = 0.0;
            const float d = (rr * rr + ri * ri <= 1e-36) ? 0.f : atan2f((float)ri, (float)rr);
            const float ax = (float)xk.x, ay = (float)xk.y;
            s_mag[gg * NBP + k] = sqrtf(ax * ax + ay * ay);
            s_del[gg * NBP + k] = d * 0.15915494309189535f;
          }
          {
            double zr = xm.x * pm[q].x + xm.y * pm[q].y, zi = xm.y * pm[q].x - xm.x * pm[q].y;
            double rr, ri;
            switch (mbin & 3) {
              case 1: rr = zi; ri = -zr; break;
              case 2: rr = -zr; ri = -zi; break;
              case 3: rr = -zi; ri = zr; break;
              default: rr = zr; ri = zi; break;
            }
            if (k == 0) ri = 0.0;
            const float d = (rr * rr + ri * ri <= 1e-36) ? 0.f : atan2f((float)ri, (float)rr);
            const float ax = (float)xm.x, ay = (float)xm.y;
            s_mag[gg * NBP + mbin] = sqrtf(ax * ax + ay * ay);
            s_del[gg * NBP + mbin] = d * 0.15915494309189535f;
          }
        }
        pk[q] = xk;
        pm[q] = xm;
      }
    }
    __syncthreads();

    // ---- gather phase: bin shift, exact phase increment, chunk-local scan, spill to HBM
#pragma unroll 1
    for (int gg = (bi == 0 ? 1 : 0); gg < G; ++gg) {
      const long long ff = f_first + gg;
      if (ff >= b) break;
      const float r = tr.rate_pf ? tr.rate_pf[ff] : wv.rate;
      const float* mg = s_mag + gg * NBP;
      const float* dl = s_del + gg * NBP;
      const size_t row = (row0 + (size_t)(ff - wv.wb)) * NBP;
#pragma unroll
      for (int q = 0; q < QB; ++q) {
        const int j = tid + q * THREADS;
        if (j >= NB) continue;
        int klo, khi;
        gather_range(j, r, NC, klo, khi);
        float smag = 0.f;
        uint32_t inc;
        if (klo <= khi) {
          for (int k = klo; k <= khi; ++k) smag += mg[k];
          // frac(r * nu / osamp) = frac(r * (khi/4 + d/(2 pi))), evaluated in double
          double tt = (double)r * (0.25 * (double)khi + (double)dl[khi]);
          tt -= floor(tt);
          inc = (uint32_t)(unsigned long long)__double2ll_rn(tt * 4294967296.0);
        } else {
          inc = ((uint32_t)j & 3u) << 30;  // s_nu = j  ->  frac(j/4)
        }
        lacc[q] += inc;
        if (ff < wv.we) tot[q] = lacc[q];
        sc.smag[row + j] = smag;
        sc.lacc[row + j] = lacc[q];
      }
    }

    // ---- peak bin (lowest k on exact ties) and f0, one warp per frame
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int gg = warp + (bi == 0 ? 1 : 0); gg < G; gg += THREADS / 32) {
        const long long ff = f_first + gg;
        if (ff >= b || ff >= wv.we) break;
        const float* mg = s_mag + gg * NBP;
        float best = -1.f;
        int bk = wv.kmin;
        for (int k = wv.kmin + lane; k <= wv.kmax; k += 32) {
          const float v = mg[k];
          if (v > best) { best = v; bk = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
          if (ov > best || (ov == best && ok < bk)) { best = ov; bk = ok; }
        }
        if (lane == 0) {
          if (tr.peak) tr.peak[ff] = bk;
          if (tr.f0) tr.f0[ff] = ((float)bk + 4.f * s_del[gg * NBP + bk]) * wv.fs_over_N;
        }
      }
    }
    __syncthreads();  // shared buffers are reused by the next batch
  }

#pragma unroll
  for (int q = 0; q < QB; ++q) {
    const int j = tid + q * THREADS;
    if (j < NB) sc.tot[trow + j] = tot[q];
  }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan over chunk totals, one thread per (track, bin)
__global__ void pv_scan_kernel(int nb, int nbp, int nchunks, const PvScratch sc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const size_t tr = blockIdx.y;
  uint32_t run = sc.carry[tr * nbp + j];
  for (int c = 0; c < nchunks; ++c) {
    const size_t i = (tr * nchunks + c) * nbp + j;
    sc.pre[i] = run;
    run += sc.tot[i];
  }
  sc.carry[tr * nbp + j] = run;
}

// ------------------------------------------------------------------------------------------------
// K_S
template <int N, int G>
__global__ void __launch_bounds__(PvCfg<N, G>::THREADS, 2)
pv_synth_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = PvCfg<N, G>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, H = Cfg::H, NBP = Cfg::NBP;
  constexpr int THREADS = Cfg::THREADS, BUF = Cfg::BUF, QP = Cfg::QP;
  using C = cplx<float>;
  using F = Fft<float, NC, +1>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                 // [G][BUF]
  float* carry = reinterpret_cast<float*>(buf + G * BUF);  // [2][3][H]

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  const PvTrack tr = tracks[blockIdx.y];
  if (tr.out == nullptr) return;
  const long long hop_lim = min(wv.we, tr.F);
  const long long a = wv.wb + (long long)blockIdx.x * wv.CS;
  if (a >= hop_lim) return;
  const long long b = min(a + (long long)wv.CS, hop_lim);
  const long long flim = min(b + 3, tr.F);  // frames [a, flim) contribute to hops [a, b)

  FftTwiddles<float, NC, +1> twd;
  twd.init(t, tb.tw_f);
  float wreg[32];
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 w2 = *reinterpret_cast<const float2*>(tb.wsyn + 2 * (t + m * TPF));
    wreg[2 * m] = w2.x;
    wreg[2 * m + 1] = w2.y;
  }
  C wr[QP];
  uint32_t prek[QP], prem[QP];
#pragma unroll
  for (int q = 0; q < QP; ++q) {
    const int k = tid + q * THREADS;
    wr[q] = (k <= NC / 2) ? tb.twr_f[k] : C{1.f, 0.f};
    prek[q] = prem[q] = 0u;
  }
  const GroupBar<TPF> bar = make_group_bar<TPF>(g, tid);

  const int nbatch = (int)((flim - a + G - 1) / G);
  const size_t row0 = (size_t)blockIdx.y * wv.rows;
  int cur_chunk = -1;
  int cb = 0;  // carry buffer holding partial sums of the three pending hops

  for (int bi = 0; bi < nbatch; ++bi) {
    const long long fb = a + (long long)bi * G;

    // ---- synthesis spectrum Y = smag e^{i theta}, folded for the N/2-point complex inverse
#pragma unroll 1
    for (int gg = 0; gg < G; ++gg) {
      const long long ff = fb + gg;
      if (ff >= flim) break;
      const int ca = (int)((ff - wv.wb) / wv.CA);
      if (ca != cur_chunk) {
        cur_chunk = ca;
        const size_t prow = ((size_t)blockIdx.y * wv.nchunksA + ca) * NBP;
#pragma unroll
        for (int q = 0; q < QP; ++q) {
          const int k = tid + q * THREADS;
          if (k > NC / 2) continue;
          prek[q] = sc.pre[prow + k];
          prem[q] = sc.pre[prow + NC - k];
        }
      }
      const size_t row = (row0 + (size_t)(ff - wv.wb)) * NBP;
      C* zb = buf + gg * BUF;
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        const int k = tid + q * THREADS;
        if (k > NC / 2) continue;
        const int mbin = NC - k;
        const float mk = sc.smag[row + k], mm = sc.smag[row + mbin];
        const uint32_t ak = prek[q] + sc.lacc[row + k], am = prem[q] + sc.lacc[row + mbin];
        float sk, ck, sm, cm;
        sincospif((float)(int)ak * 4.656612873077393e-10f, &sk, &ck);
        sincospif((float)(int)am * 4.656612873077393e-10f, &sm, &cm);
        if (k == 0) {
          const float y0 = mk * ck, yn = mm * cm;  // Im of DC / Nyquist forced to 0
          zb[fft_pad(0)] = C{y0 + yn, y0 - yn};
        } else {
          const float ykr = mk * ck, yki = mk * sk, ymr = mm * cm, ymi = mm * sm;
          // A = Y_k, B = conj(Y_m): E2 = A + B, D2 = A - B, O2 = D2 * conj(W^k)
          const float er = ykr + ymr, ei = yki - ymi;
          const float dr = ykr - ymr, di = yki + ymi;
          const float orr = dr * wr[q].x + di * wr[q].y, oi = di * wr[q].x - dr * wr[q].y;
          zb[fft_pad(k)] = C{er - oi, ei + orr};
          if (mbin != k) zb[fft_pad(mbin)] = C{er + oi, orr - ei};
        }
      }
    }
    __syncthreads();

    // ---- inverse FFT, synthesis window (includes gain and 1/N), result in place as real pairs
    {
      const long long fg = fb + g;
      if (fg < flim) {
        C x[16];
        C* zb = buf + g * BUF;
        F::load(x, zb, t);
        bar.sync();
        F::run(x, zb, t, twd, bar);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          x[m].x *= wreg[2 * m];
          x[m].y *= wreg[2 * m + 1];
        }
        F::store(x, zb, t);
      }
    }
    __syncthreads();

    // ---- overlap-add.  Hop h = samples [hH, (h+1)H) = sum over frames f = h..h+3 of
    //      y_f[(h-f+3)H + i], added in ascending f.  This batch completes hops fb-3 .. fb+G-4 and
    //      leaves partial sums of the last three in the other carry buffer.
    {
      const float* cin = carry + cb * 3 * H;
      float* cout = carry + (cb ^ 1) * 3 * H;
      const long long last_f = min(fb + G, flim) - 1;
      for (int it = tid; it < (G + 3) * (H / 2); it += THREADS) {
        const int hh = it / (H / 2), i2 = it % (H / 2);
        const long long h = fb - 3 + hh;
        if (h < a) continue;  // hops before the chunk belong to the previous CTA
        float2 s = make_float2(0.f, 0.f);
        if (hh < 3) s = *reinterpret_cast<const float2*>(cin + hh * H + 2 * i2);
        const long long f_lo = max(h, fb), f_hi = min(h + 3, last_f);
        for (long long f = f_lo; f <= f_hi; ++f) {
          const int cidx = (int)(h - f + 3) * (H / 2) + i2;  // complex index inside frame f
          const C v = buf[(int)(f - fb) * BUF + fft_pad(cidx)];
          s.x += v.x;
          s.y += v.y;
        }
        const bool complete = min(h + 3, flim - 1) <= last_f;
        if (complete) {
          if (h < b) {
            const long long o = h * H + 2 * i2;
            if (o + 1 < tr.n) {
              *reinterpret_cast<float2*>(tr.out + o) = s;
            } else if (o < tr.n) {
              tr.out[o] = s.x;
            }
          }
        } else {
          *reinterpret_cast<float2*>(cout + (hh - G) * H + 2 * i2) = s;
        }
      }
      cb ^= 1;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
template <int N>
static cudaError_t configure_n() {
  constexpr int G = PvG<N>::value;
  cudaError_t e = cudaFuncSetAttribute(pv_analyze_kernel<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)PvCfg<N, G>::SMEM_A);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pv_synth_kernel<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)PvCfg<N, G>::SMEM_S);
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
int pv_threads(int) { return 256; }
size_t pv_analyze_smem(int fftN) {
  switch (fftN) {
    case 512: return PvCfg<512, 16>::SMEM_A;
    case 1024: return PvCfg<1024, 8>::SMEM_A;
    case 2048: return PvCfg<2048, 4>::SMEM_A;
    case 4096: return PvCfg<4096, 2>::SMEM_A;
    case 8192: return PvCfg<8192, 1>::SMEM_A;
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
                              const PvTables& tb, const PvScratch& sc, cudaStream_t st) {
  MLX_PV_DISPATCH(fftN, {
    constexpr int G = PvG<N>::value;
    dim3 grid(wv.nchunksA, ntracks);
    pv_analyze_kernel<N, G><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_A, st>>>(tracks, wv, tb, sc);
  });
  return cudaGetLastError();
}

cudaError_t launch_pv_scan(int fftN, int ntracks, const PvWave& wv, const PvScratch& sc, cudaStream_t st) {
  const int nb = fftN / 2 + 1;
  dim3 grid((nb + 255) / 256, ntracks);
  pv_scan_kernel<<<grid, 256, 0, st>>>(nb, pv_nbp(fftN), wv.nchunksA, sc);
  return cudaGetLastError();
}

cudaError_t launch_pv_synth(int fftN, const PvTrack* tracks, int ntracks, const PvWave& wv,
                            const PvTables& tb, const PvScratch& sc, cudaStream_t st) {
  MLX_PV_DISPATCH(fftN, {
    constexpr int G = PvG<N>::value;
    dim3 grid(wv.nchunksS, ntracks);
    pv_synth_kernel<N, G><<<grid, PvCfg<N, G>::THREADS, PvCfg<N, G>::SMEM_S, st>>>(tracks, wv, tb, sc);
  });
  return cudaGetLastError();
}

}  // namespace mlx
