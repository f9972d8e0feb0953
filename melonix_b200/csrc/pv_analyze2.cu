// melonix_b200/csrc/pv_analyze2.cu -- K_A2: the phase-vocoder analysis kernel for a constant pitch ratio
// >= 1 (PV-spec v1; NOT IN REFERENCE -- melonix has no phase vocoder).
// Same results, bit for bit, as the general kernel pv_analyze_kernel (pv_kernels.cu), which stays the
// path for ratios < 1, per-frame ratios and the other transform sizes.
//
// What is different, and why (round-1 ncu: the shared-memory data pipe was the busiest unit of K_A, 62 %,
// with 23 % of its wavefronts bank conflicts; 49 % issue-slot utilisation behind three CTA-wide barriers
// per batch):
//
//  * The LAST FFT stage moves from the FFT threads to the threads that own the bins.  The Stockham plan of
//    the N/2-point transform is 16 x 16 x R (R = 2, 4, 8 for fftN = 1024, 2048, 4096): the frame's FFT group
//    runs the two radix-16 stages and leaves 256-point sub-transforms in shared memory; the radix-R
//    butterfly j produces bins j + 256 d, its mirror butterfly 256 - j produces exactly the mirrored bins
//    N/2 - (j + 256 d).  Two lanes of one warp (lane, lane ^ 16) take the two butterflies and swap half of
//    their outputs with warp shuffles, after which each holds R/2 complete (k, N/2 - k) pairs in registers.
//    Gone: the last-stage load, the natural-order store, the pair-phase loads (3 x 16 KB per frame of
//    16-byte shared-memory traffic at fftN = 2048) and one group barrier.
//  * Bin shift as a SCATTER from the thread that analysed the bin.  For a ratio >= 1 the map
//    k -> j = trunc(float(k) * r) is injective, so the thread that owns input bin k owns output bin j (and
//    the empty output bins up to the next one) for the whole launch: the exact phase increment, the running
//    phase and the two global stores happen right where magnitude and phase advance were computed.  Gone:
//    the (mag, d) records in shared memory (8-byte accesses at a 16-byte stride: all of the kernel's bank
//    conflicts), the gather phase and its CTA-wide barrier.  Input bins whose output bin lies beyond the
//    Nyquist bin are not analysed at all (16 % of the bins at +3 semitones).
//  * The stage-1 output is stored LINEARLY (stage-1 stores and the tail loads are unit-stride, for the
//    mirrored butterflies at any alignment): conflict-free without padding.  Only the stage-0 exchange
//    keeps the padded layout of fft.cuh.
//  * One TMA tile buffer instead of two: the refill is issued as soon as the FFT groups have taken their
//    samples out and lands under the (long) bin phase.
//
// Per batch of G frames: [TMA wait] FFT stages 0-1 | __syncthreads | tail butterflies, pair split,
// magnitude / integer-turn phase / FP64 cut decision, scatter | __syncthreads | peak bin + f0 (one warp per
// frame).  Two CTA-wide barriers (three before).
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft.cuh"
#include "kernels.h"
#include "pv_analysis.cuh"
#include "pv_common.cuh"
#include "pv_shift.cuh"
#include "tma.cuh"

#ifndef MLX_KA2_CTAS
#define MLX_KA2_CTAS 2
#endif
#ifndef MLX_KA2_UNROLL
#define MLX_KA2_UNROLL 2  // frames of the bin phase in flight per thread
#endif

namespace mlx {

constexpr int kKa2Unroll = MLX_KA2_UNROLL;

template <int N>
struct Ka2Cfg {
  static constexpr int NC = N / 2;
  using P = FftPlan<NC>;
  static_assert(P::A == 2 && P::B >= 1, "K_A2 needs a 16 x 16 x R plan (fftN = 1024, 2048, 4096)");
  static constexpr int R = 1 << P::B;      // radix of the tail stage
  static constexpr int NQ = NC / R;        // tail butterflies per frame
  static_assert(NQ == 256, "two radix-16 stages leave 256-point sub-transforms");
  static constexpr int THREADS = 256;      // = 2 lanes per (j, 256 - j) unit, j < 128
  static constexpr int TPF = NC / 16;      // FFT threads per frame
  static constexpr int G = THREADS / TPF;  // frames per batch
  static constexpr int H = N / 4;
  static constexpr int NB = NC + 1, NBP = NC + 32;
  static constexpr int BUF = P::BUF;       // padded complex slots per frame (stage-0 exchange)
  static constexpr int TILE = N + (G - 1) * H;
  static constexpr int SLOTS = R / 2;      // (k, N/2 - k) pairs per thread
  static constexpr int PKB = N / 16;       // capacity of the peak-search band [kmin, kmax]
  static constexpr bool WIN_D = (N <= 2048);
  static constexpr size_t SMEM = sizeof(cplx<double>) * G * BUF + (WIN_D ? sizeof(double) * N : 0) +
                                 sizeof(float) * TILE + 8 * G * PKB + 16;
};

// the FFT role's only twiddle register (shape expected by Fft<>::compute)
struct Ka2Twiddle {
  static constexpr bool kPre1 = false;
  cplx<double> w[1];
};

// what the scatter needs to know about one input bin, decoded from PvWave::dst[k]
struct BinDst {
  int j;        // output bin, -1: beyond the Nyquist bin (the input bin is dropped)
  int nz;       // empty output bins j+1 .. j+nz that follow (they belong to this bin's owner)
  unsigned long long base;  // r_fix * k * 2^30 + 2^25 (pv_shift.cuh)
};
__device__ __forceinline__ BinDst make_bin_dst(const uint32_t* __restrict__ dst, int k, uint32_t r_fix) {
  const uint32_t e = __ldg(dst + k);
  BinDst b;
  b.j = (e == 0xffffffffu) ? -1 : (int)(e & 0xffffu);
  b.nz = (e == 0xffffffffu) ? 0 : (int)(e >> 16);
  b.base = (unsigned long long)r_fix * ((unsigned long long)k << 30) + (1ULL << 25);
  return b;
}

// running state of one analysed bin across the frames of a chunk
struct BinState {
  cplx<double> x;   // previous frame's spectrum value (for the FP64 decision at the +-pi cut)
  uint32_t p;       // previous frame's phase, integer turns
  float mag;        // previous frame's magnitude (silence gate)
  uint32_t lacc;    // chunk-local running synthesis phase of the output bin
};

template <int N>
__global__ void __launch_bounds__(Ka2Cfg<N>::THREADS, MLX_KA2_CTAS)
pv_analyze2_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = Ka2Cfg<N>;
  using P = typename Cfg::P;
  constexpr int NC = Cfg::NC, R = Cfg::R, NQ = Cfg::NQ, TPF = Cfg::TPF, G = Cfg::G, H = Cfg::H;
  constexpr int NB = Cfg::NB, NBP = Cfg::NBP, BUF = Cfg::BUF, TILE = Cfg::TILE, SLOTS = Cfg::SLOTS;
  constexpr int THREADS = Cfg::THREADS, PKB = Cfg::PKB;
  constexpr bool WD = Cfg::WIN_D;
  using C = cplx<double>;
  using F = Fft<double, NC, -1>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                        // [G][BUF]
  double* s_win = reinterpret_cast<double*>(buf + G * BUF);       // [N] when WD
  float* tile = reinterpret_cast<float*>(s_win + (WD ? N : 0));   // [TILE]
  float* pk_mag = tile + TILE;                                    // [G][PKB] magnitudes of the search band (sign = cut flip)
  int* pk_d = reinterpret_cast<int*>(pk_mag + G * PKB);           // [G][PKB] their wrapped phase advance
  uint64_t* mbar = reinterpret_cast<uint64_t*>(pk_d + G * PKB);

  const int tid = threadIdx.x;
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

  if (tid == 0) mbar_init(mbar, 1);
  if constexpr (WD) {
    for (int i = tid; i < N; i += THREADS) s_win[i] = tb.win_d[i];
  }
  __syncthreads();  // mbarrier initialised, window staged
  if (tid == 0) {   // first tile: samples [(a-1-3)H, (a-1+G)H) of the zero-padded track
    mbar_expect_tx(mbar, TILE * sizeof(float));
    tma_load_1d(tile, tr.x + (a - 4) * H, TILE * sizeof(float), mbar);
  }

  // ---- FFT role: thread t of the group that transforms frame f_first + g
  const int g = tid / TPF, t = tid % TPF;
  Ka2Twiddle tw1;  // twiddle of the second radix-16 stage: exp(-2 pi i (t mod 16) / 256)
  tw1.w[0] = tb.tw_d[(t & 15) * (NC / 256)];
  const GroupBar<TPF> bar = make_group_bar<TPF>(g, tid);

  // ---- bin role: lanes (l, l ^ 16) of a warp share the butterfly pair (j, 256 - j)
  const int warp = tid >> 5, lane = tid & 31, hside = lane >> 4;
  const int ju = warp * 16 + (lane & 15);             // unit, 0 .. 127
  const bool special = (ju == 0);                     // butterflies 0 and 128 mirror onto themselves
  const int bj = special ? (hside ? NQ / 2 : 0) : (hside ? NQ - ju : ju);
  const C wtail = tb.tw_d[bj];                        // exp(-2 pi i bj / NC): twiddle of tail butterfly bj
  const uint32_t r_fix = (uint32_t)wv.r_fix;
  // slot s: bins kS = bj + 256 s and NC - kS.  Thread 0 (bj = 0): slot 0 = the real bins 0 and NC, slots
  // s >= 1 = (256 s, NC - 256 s) from its own outputs, plus the self-mirrored bin NC/2 (extra state).
  C wpair[SLOTS];
  BinState sk[SLOTS], sm[SLOTS];
  BinDst dk[SLOTS], dm[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int k = bj + NQ * s;
    wpair[s] = tb.twr_d[k];  // exp(-2 pi i k / N), k <= NC/2
    dk[s] = make_bin_dst(wv.dst, k, r_fix);
    dm[s] = make_bin_dst(wv.dst, NC - k, r_fix);
    sk[s] = BinState{C{1.0, 0.0}, 0u, 1.f, 0u};  // frame -1 has phi = 0 <=> X = 1
    sm[s] = sk[s];
  }
  BinState sself{C{1.0, 0.0}, 0u, 1.f, 0u};  // bin NC/2 (thread 0 only)
  const BinDst dself = make_bin_dst(wv.dst, NC / 2, r_fix);
  const int kmin = wv.kmin, kmax = wv.kmax;
  const size_t row0 = (size_t)blockIdx.y * wv.rows;
  uint32_t emitted = 0u;  // frames of this chunk that have emitted so far (empty output bins advance by j/4 turn each)

  // one analysed bin -> its output bin j (+ the empty bins that follow): exact phase increment, running
  // phase, two global stores.  `mb` = magnitude bits with the cut-flip flag in the sign bit.
  auto scatter = [&](const BinDst& d, BinState& st, float mag, int dq, bool flip, float* rs, uint32_t* rl) {
    if (d.j < 0) return;
    const uint32_t mb = __float_as_uint(mag) | (flip ? 0x80000000u : 0u);
    st.lacc += shift_inc(d.base, dq, mb, (int)r_fix);
    rs[d.j] = mag;
    rl[d.j] = st.lacc;
    for (int z = 1; z <= d.nz; ++z) {  // empty K_j: smag = 0, s_nu = j -> inc = frac(j / 4) per frame
      rs[d.j + z] = 0.f;
      rl[d.j + z] = (emitted * (uint32_t)((d.j + z) & 3)) << 30;
    }
  };
  // phase totals of the chunk's output bins (`cnt` frames emitted): tot / totc rows
  auto put_total = [&](uint32_t* row, const BinDst& d, const BinState& st, uint32_t cnt) {
    if (d.j < 0) return;
    row[d.j] = st.lacc;
    for (int z = 1; z <= d.nz; ++z) row[d.j + z] = (cnt * (uint32_t)((d.j + z) & 3)) << 30;
  };
  auto put_totals = [&](uint32_t* row, uint32_t cnt) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      put_total(row, dk[s], sk[s], cnt);
      if (!(special && !hside && s == 0) || true) put_total(row, dm[s], sm[s], cnt);
    }
    if (tid == 0) put_total(row, dself, sself, cnt);
  };

  for (int bi = 0; bi < nbatch; ++bi) {
    const long long f_first = a - 1 + (long long)bi * G;
    mbar_wait(mbar, bi & 1);

    // ---- FFT role: window, radix-16 stage 0 -> padded exchange -> radix-16 stage 1 -> linear store
    {
      const long long fg = f_first + g;
      if (fg >= 0 && fg < b) {
        C x[16];
        const float* src = tile + g * H;
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
        C* fb = buf + g * BUF;
        F::template compute<0>(x, fb, t, tw1);
        bar.sync();
        F::load(x, fb, t);
        bar.sync();  // every load of the group done before the linear stores overwrite the padded slots
        {
          C v[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) v[r] = x[r];
          twiddle_powers<16>(v, tw1.w[0]);
          dft16<-1>(v);
          const int k = t & 15;
          C* p = fb + (t - k) * 16 + k;  // stage-1 butterfly t writes elements (t-k)*16 + k + 16 r
#pragma unroll
          for (int r = 0; r < 16; ++r) p[r * 16] = v[r];
        }
      }
    }
    __syncthreads();  // sub-transforms of all frames in place; the tile has been consumed
    if (tid == 0 && bi + 1 < nbatch) {
      fence_proxy_async();  // generic-proxy reads of the tile (ordered by the barrier) before the async-proxy refill
      mbar_expect_tx(mbar, TILE * sizeof(float));
      tma_load_1d(tile, tr.x + (f_first + G - 3) * H, TILE * sizeof(float), mbar);
    }

    // ---- bin role
    const int g_hi = (int)min((long long)G, b - f_first);  // frames of this batch that exist: [g_lo, g_hi)
    const int g_lo = f_first < 0 ? 1 : 0;
#pragma unroll(kKa2Unroll)
    for (int gg = 0; gg < G; ++gg) {
      if (gg >= g_hi) break;
      if (gg < g_lo) continue;
      const long long ff = f_first + gg;
      const bool emit = (bi != 0 || gg != 0);  // the chunk's leading halo frame only seeds the state
      const C* zb = buf + gg * BUF + bj;
      C v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = zb[r * NQ];
      twiddle_powers<R>(v, wtail);
      dft_r<R, -1>(v);  // v[d] = Z[bj + 256 d]
      // mirrored partners: slot s pairs Z[bj + 256 s] with the other lane's Z[(256 - bj) + 256 (R-1-s)]
      C zc[SLOTS];
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const C mine = v[R - 1 - s];
        C got;
        got.x = __shfl_xor_sync(0xffffffffu, mine.x, 16);
        got.y = __shfl_xor_sync(0xffffffffu, mine.y, 16);
        // the self-mirrored butterflies: 128 pairs its outputs d <-> R-1-d, 0 pairs d <-> R-d
        zc[s] = special ? (hside ? mine : v[(R - s) % R]) : got;
      }
      if (emit) ++emitted;
      float* rs = sc.smag + (row0 + (size_t)(ff - wv.wb)) * NBP;
      uint32_t* rl = sc.lacc + (row0 + (size_t)(ff - wv.wb)) * NBP;
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        const int k = bj + NQ * s, mbin = NC - k;
        if (s == 0 && tid == 0) {
          // DC and Nyquist are real: X[0] = Re Z0 + Im Z0, X[NC] = Re Z0 - Im Z0
          const C z0 = v[0];
          const MagD m0 = analysis_real_bin(z0.x + z0.y, sk[0].p, sk[0].mag);
          const MagD mn = analysis_real_bin(z0.x - z0.y, sm[0].p, sm[0].mag);
          if (emit) {
            scatter(dk[0], sk[0], fabsf(m0.mag), m0.d, __float_as_uint(m0.mag) >> 31, rs, rl);
            scatter(dm[0], sm[0], fabsf(mn.mag), mn.d, __float_as_uint(mn.mag) >> 31, rs, rl);
          }
          continue;
        }
        const bool need_k = dk[s].j >= 0, need_m = dm[s].j >= 0;
        if (!need_k && !need_m) continue;
        const C za = v[s], zz = zc[s], w = wpair[s];
        const double er = 0.5 * (za.x + zz.x), ei = 0.5 * (za.y - zz.y);
        const double dr = 0.5 * (za.x - zz.x), di = 0.5 * (za.y + zz.y);
        const double tr_ = dr * w.x - di * w.y, ti_ = dr * w.y + di * w.x;
        float mag;
        int dq;
        bool flip;
        if (need_k) {
          const C xk{er + ti_, ei - tr_};
          analysis_bin(xk.x, xk.y, sk[s].x.x, sk[s].x.y, sk[s].p, sk[s].mag, k, false, mag, dq, flip);
          sk[s].x = xk;
          if (emit) {
            scatter(dk[s], sk[s], mag, dq, flip, rs, rl);
            if (s == 0 && k >= kmin && k <= kmax) {
              pk_mag[gg * PKB + k - kmin] = flip ? -mag : mag;
              pk_d[gg * PKB + k - kmin] = dq;
            }
          }
        }
        if (need_m) {
          const C xm{er - ti_, -ei - tr_};
          analysis_bin(xm.x, xm.y, sm[s].x.x, sm[s].x.y, sm[s].p, sm[s].mag, mbin, false, mag, dq, flip);
          sm[s].x = xm;
          if (emit) scatter(dm[s], sm[s], mag, dq, flip, rs, rl);
        }
      }
      if (tid == 0 && dself.j >= 0) {  // bin NC/2 mirrors onto itself: X = conj(Z[NC/2])
        const C zs = v[R / 2];
        float mag;
        int dq;
        bool flip;
        analysis_bin(zs.x, -zs.y, sself.x.x, sself.x.y, sself.p, sself.mag, NC / 2, false, mag, dq, flip);
        sself.x = C{zs.x, -zs.y};
        if (emit) scatter(dself, sself, mag, dq, flip, rs, rl);
      }
      if (emit && ff == wv.we - 1) put_totals(sc.totc + trow, emitted);  // phase the next wave starts from
    }
    __syncthreads();  // frame buffers free for the next batch; the band records of this one complete

    // ---- peak bin (lowest k on exact ties) and f0, one warp per frame
    for (int gg = warp + (bi == 0 ? 1 : 0); gg < G; gg += THREADS / 32) {
      const long long ff = f_first + gg;
      if (ff >= b || ff >= wv.we) break;
      const float* pm = pk_mag + gg * PKB - kmin;
      float best = -1.f;
      int bk = kmin;
      for (int k = kmin + lane; k <= kmax; k += 32) {
        const float vv = fabsf(pm[k]);
        if (vv > best) { best = vv; bk = k; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ov > best || (ov == best && ok < bk)) { best = ov; bk = ok; }
      }
      if (lane == 0) {
        if (tr.peak) tr.peak[ff] = bk;
        if (tr.f0) {
          long long dd = (long long)pk_d[gg * PKB + bk - kmin];
          if (__float_as_uint(pm[bk]) >> 31) dd += (dd < 0) ? 4294967296LL : -4294967296LL;
          tr.f0[ff] = ((float)bk + (float)dd * 9.313225746154785e-10f) * wv.fs_over_N;  // nu = k + 4 d
        }
      }
    }
    // (the next batch writes the band records only after its first __syncthreads)
  }

  put_totals(sc.tot + trow, emitted);  // all frames of the chunk: prefix of the later chunks of this wave
  if (b <= wv.we) {
    put_totals(sc.totc + trow, emitted);  // every frame of the chunk lies before the wave end
  } else if (a >= wv.we) {
    for (int j = tid; j < NB; j += THREADS) sc.totc[trow + j] = 0u;  // none does
  }
}

// ------------------------------------------------------------------------------------------------
bool pv_analyze2_supported(int fftN) { return fftN == 1024 || fftN == 2048; }
int pv_analyze2_band_capacity(int fftN) { return fftN / 16; }

template <int N>
static cudaError_t configure2_n() {
  cudaError_t e = cudaFuncSetAttribute(pv_analyze2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Ka2Cfg<N>::SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pv_analyze2_kernel<N>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              (int)cudaSharedmemCarveoutMaxShared);
}

cudaError_t pv_analyze2_configure(int fftN) {
  switch (fftN) {
    case 1024: return configure2_n<1024>();
    case 2048: return configure2_n<2048>();
    default: return cudaSuccess;
  }
}
int pv_analyze2_frames_per_batch(int fftN) {
  switch (fftN) {
    case 1024: return Ka2Cfg<1024>::G;
    case 2048: return Ka2Cfg<2048>::G;
    default: return 0;
  }
}

cudaError_t launch_pv_analyze2(int fftN, const PvTrack* tracks, int ntracks, const PvWave& wv, const PvTables& tb,
                               const PvScratch& sc, cudaStream_t st) {
  dim3 grid(wv.nchunksA, ntracks);
  switch (fftN) {
    case 1024:
      pv_analyze2_kernel<1024><<<grid, Ka2Cfg<1024>::THREADS, Ka2Cfg<1024>::SMEM, st>>>(tracks, wv, tb, sc);
      break;
    case 2048:
      pv_analyze2_kernel<2048><<<grid, Ka2Cfg<2048>::THREADS, Ka2Cfg<2048>::SMEM, st>>>(tracks, wv, tb, sc);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace mlx
