// melonix_b200/csrc/pv_analyze2.cu -- K_A2: the phase-vocoder analysis kernel for a constant pitch ratio
// >= 1 (PV-spec v1; NOT IN REFERENCE -- melonix has no phase vocoder).
// Same results, bit for bit, as the general kernel pv_analyze_kernel (pv_kernels.cu), which stays the
// path for ratios < 1, per-frame ratios and the other transform sizes.
//
// What is different, and why (round-1 ncu: the shared-memory data pipe was the busiest unit of K_A, 62 %,
// with 23 % of its wavefronts bank conflicts):
//
//  * The LAST FFT stage moves from the FFT threads to the threads that own the bins.  The Stockham plan of
//    the N/2-point transform is 16 x 16 x R (R = 2, 4 for fftN = 1024, 2048): the frame's FFT group runs the
//    two radix-16 stages and leaves 256-point sub-transforms in shared memory; the radix-R butterfly j
//    produces bins j + 256 d, its mirror butterfly 256 - j produces exactly the mirrored bins
//    N/2 - (j + 256 d).  Two lanes of one warp (lane, lane ^ 16) take the two butterflies and swap half of
//    their outputs with warp shuffles, after which each holds R/2 complete (k, N/2 - k) pairs in registers.
//    Gone: the last-stage load, the natural-order store, the pair-phase loads (3 x 16 KB per frame of
//    16-byte shared-memory traffic at fftN = 2048) and one group barrier.
//  * The (magnitude, phase advance) records of a pair go back, as ONE 16-byte store, into the slot the
//    thread itself loaded its butterfly input from (nobody else reads that slot): record of bin k < N/4 in
//    the first half of slot k, of bin N/2 - k in the second half.  No 8-byte stores at a 16-byte stride (the
//    general kernel's bank conflicts), no extra shared memory.
//  * The stage-1 output is stored LINEARLY (stage-1 stores and the tail loads are unit-stride, for the
//    mirrored butterflies at any alignment): conflict-free without padding.  Only the stage-0 exchange
//    keeps the padded layout of fft.cuh.
//  * Bins 0, N/2 (real) and N/4 (self-mirrored) are left over by the pairing; analysing them in thread 0
//    would make warp 0 the straggler of every bin phase.  Thread 0 parks Z[0] and Z[N/4] in shared memory
//    and a warp that idles after the barrier takes them, one lane per frame, writing their output bins
//    directly (for a ratio >= 1 every output bin is fed by at most one input bin).
//  * One TMA tile buffer instead of two: the refill is issued as soon as the FFT groups have taken their
//    samples out and lands under the (long) bin phase.
//  * The phase increment uses inc = A + ((d r + 2^25) >> 26) + flip term with the 32-bit per-bin constant
//    A = 16 k r mod 2^32 (the same value as pv_shift.cuh's 64-bit base form: r k 2^30 is a multiple of 2^26).
//
// Per batch of G frames: [TMA wait] FFT stages 0-1 | __syncthreads | tail butterflies, pair split,
// magnitude / integer-turn phase / FP64 cut decision, records | __syncthreads | bin shift (gather) with
// coalesced 8-byte stores, Nyquist output bin, peak bin + f0, left-over bins | __syncthreads.
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft.cuh"
#include "kernels.h"
#include "pv_analysis.cuh"
#include "pv_common.cuh"
#ifndef MLX_KA_TAB1
#define MLX_KA_TAB1 0
#endif
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
  static constexpr int QB = NC / THREADS;  // output bins per thread in the gather (bin NC: last warp)
  static constexpr int ZSLOT = BUF - 1;    // an all-zero record (what an empty K_j reads); the padded exchange ends at BUF - 2
  static constexpr bool WIN_D = (N <= 2048);
  static constexpr size_t SMEM = sizeof(cplx<double>) * G * BUF + (WIN_D ? sizeof(double) * N : 0) +
                                 sizeof(float) * TILE + sizeof(cplx<double>) * 2 * G + 16;
};

// the FFT role's only twiddle register (shape expected by Fft<>::compute)
struct Ka2Twiddle {
  static constexpr bool kPre1 = false;
  static constexpr bool kTab1 = false;  // (this kernel applies the stage-1 table itself)
  cplx<double> w[1];
};

// running state of one analysed bin across the frames of a chunk
struct BinState {
  cplx<double> x;   // previous frame's spectrum value (for the FP64 decision at the +-pi cut)
  uint32_t p;       // previous frame's phase, integer turns
  float mag;        // previous frame's magnitude (silence gate)
};

// exp(-2 pi i (bj + 256 s) / N) from w0 = exp(-2 pi i bj / N): a rotation by the constant angle
// -2 pi 256 s / N (N = 2048: s = 1 -> -pi/4; N = 4096: -pi/8 steps), 4 FP64 operations instead of registers
// trunc(float(k) * r) with a plain float multiply, exactly as the spec (A.5) and the oracle do
__device__ __forceinline__ int shift_bin_ka2(int k, float r) { return (int)truncf(__fmul_rn((float)k, r)); }

template <int N>
__device__ __forceinline__ cplx<double> pair_twiddle(const cplx<double> w0, int s) {
  if (s == 0) return w0;
  // cmul_w16<-1, M>: multiply by exp(-2 pi i M / 16); 256 s / N turns = (4096 s / N) sixteenths
  constexpr int STEP = 4096 / N;  // N = 2048 -> 2, N = 4096 -> 1 (N = 1024 has one slot)
  switch (s * STEP) {
    case 1: return rot16_pinned<-1, 1>(w0);
    case 2: return rot16_pinned<-1, 2>(w0);
    case 3: return rot16_pinned<-1, 3>(w0);
    default: return w0;
  }
}

template <int N>
__global__ void __launch_bounds__(Ka2Cfg<N>::THREADS, MLX_KA2_CTAS)
pv_analyze2_kernel(const PvTrack* __restrict__ tracks, const PvWave wv, const PvTables tb, const PvScratch sc) {
  using Cfg = Ka2Cfg<N>;
  constexpr int NC = Cfg::NC, R = Cfg::R, NQ = Cfg::NQ, TPF = Cfg::TPF, G = Cfg::G, H = Cfg::H;
  constexpr int NB = Cfg::NB, NBP = Cfg::NBP, BUF = Cfg::BUF, TILE = Cfg::TILE, SLOTS = Cfg::SLOTS;
  constexpr int THREADS = Cfg::THREADS, QB = Cfg::QB, ZSLOT = Cfg::ZSLOT;
  constexpr bool WD = Cfg::WIN_D;
  using C = cplx<double>;
  using F = Fft<double, NC, -1>;
  static_assert(G <= 8, "the left-over bins use one lane per frame of an 8-lane subgroup");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* buf = reinterpret_cast<C*>(smem_raw);                        // [G][BUF]
  double* s_win = reinterpret_cast<double*>(buf + G * BUF);       // [N] when WD
  float* tile = reinterpret_cast<float*>(s_win + (WD ? N : 0));   // [TILE]
  C* s_sp = reinterpret_cast<C*>(tile + TILE);                    // [G][2] Z[0] and Z[NC/2] of each frame
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_sp + 2 * G);

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
  if (tid < G) buf[tid * BUF + ZSLOT] = C{0.0, 0.0};  // the zero record of every frame buffer (never overwritten)
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

  // ---- pair role: lanes (l, l ^ 16) of a warp share the butterfly pair (j, 256 - j)
  const int warp = tid >> 5, lane = tid & 31, hside = lane >> 4;
  const int ju = warp * 16 + (lane & 15);             // unit, 0 .. 127
  const bool special = (ju == 0);                     // butterflies 0 and 128 mirror onto themselves
  const int bj = special ? (hside ? NQ / 2 : 0) : (hside ? NQ - ju : ju);
  // exp(-2 pi i bj / NC), the twiddle of tail butterfly bj -- formed exactly as the general kernel's last FFT
  // stage forms it (fft.cuh, MLX_FFT_DERIVE_LAST: table value of bj mod TPF, rotated by (bj / TPF) sixteenths),
  // so that the two kernels stay bit-identical
  C wtail = tb.tw_d[bj % TPF];
  switch (bj / TPF) {
    case 1: wtail = rot16_pinned<-1, 1>(wtail); break;
    case 2: wtail = rot16_pinned<-1, 2>(wtail); break;
    case 3: wtail = rot16_pinned<-1, 3>(wtail); break;
    case 4: wtail = rot16_pinned<-1, 4>(wtail); break;
    case 5: wtail = rot16_pinned<-1, 5>(wtail); break;
    case 6: wtail = rot16_pinned<-1, 6>(wtail); break;
    case 7: wtail = rot16_pinned<-1, 7>(wtail); break;
    default: break;
  }
  const int r_fix = (int)wv.r_fix;
  // slot s: bins kS = bj + 256 s < NC/2 and NC - kS.  Thread 0 (bj = 0): its slot 0 would be the real bins
  // 0 and NC, and bin NC/2 (from Z[NC/2], its output R/2) pairs with itself: those three are the left-over
  // bins (see the header); slots s >= 1 of thread 0 pair its own outputs s and R - s.
  const C wpair0 = tb.twr_d[bj];  // exp(-2 pi i bj / N); slot s uses exp(-2 pi i (bj + 256 s) / N) = wpair0 * rot_s
  BinState sk[SLOTS], sm[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    sk[s] = BinState{C{1.0, 0.0}, 0u, 1.f};  // frame -1 has phi = 0 <=> X = 1
    sm[s] = sk[s];
  }

  // ---- left-over bins: warp `spw`, 8-lane subgroup = bin (NC/2, 0, NC), lane within it = frame
  const int spw = G % (THREADS / 32);
  const int spk = (lane >> 3) == 0 ? NC / 2 : ((lane >> 3) == 1 ? 0 : NC);
  const bool sp_lane = warp == spw && lane < 24;
  const bool sp_owner = sp_lane && (lane & 7) == 0;  // the lane that reports the subgroup's totals
  const int spj = shift_bin_ka2(spk, wv.rate);        // the output bin this input bin feeds (if <= NC)
  const bool sp_fed = sp_lane && spj <= NC;
  const uint32_t spA = ((uint32_t)spk * (uint32_t)r_fix) << 4;
  BinState ssp{C{1.0, 0.0}, 0u, 1.f};
  uint32_t sp_lacc = 0u;
  // exp(-2 pi i (NC/2) / N) as the general kernel forms it for its pair k = NC/2 (thread (k-1) % T, pair (k-1) / T
  // of its T threads: table value rotated by (k-1)/T * 16 T / N sixteenths; pv_kernels.cu, MLX_KA_DERIVE_WPAIR)
  C wself;
  {
    constexpr int T1 = PvCfg<N, PvG<N>::analyze>::THREADS;
    constexpr int q = (NC / 2 - 1) / T1;
    wself = tb.twr_d[NC / 2 - q * T1];
    switch (q * (16 * T1 / N)) {
      case 1: wself = rot16_pinned<-1, 1>(wself); break;
      case 2: wself = rot16_pinned<-1, 2>(wself); break;
      case 3: wself = rot16_pinned<-1, 3>(wself); break;
      case 4: wself = rot16_pinned<-1, 4>(wself); break;
      case 6: wself = rot16_pinned<-1, 6>(wself); break;
      default: break;
    }
  }

  // ---- gather role: output bins j = tid + 256 q (and bin NC: the last warp, one lane per frame)
  //   goff: byte offset of the source record inside a frame buffer (0xffffffff: fed by a left-over bin,
  //   written by the `spw` lanes);  gA: the per-bin constant of shift_inc_a
  auto source_of = [&](int j, uint32_t& off, uint32_t& A) {
    const uint32_t kk = __ldg(wv.gk + j);
    const int klo = (int)(kk & 0xffffu), kh = (int)(kk >> 16);
    if (klo > kh) {  // empty K_j: smag = 0, s_nu = j -> inc = frac(j / 4) turn per frame
      off = 16u * ZSLOT;
      A = ((uint32_t)j & 3u) << 30;
    } else if (kh == 0 || kh == NC / 2 || kh == NC) {
      off = 0xffffffffu;
      A = 0u;
    } else {
      off = kh < NC / 2 ? 16u * kh : 16u * (NC - kh) + 8u;
      A = ((uint32_t)kh * (uint32_t)r_fix) << 4;
    }
  };
  uint32_t goff[QB], gA[QB], lacc[QB], totc[QB];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    source_of(tid + q * THREADS, goff[q], gA[q]);
    lacc[q] = totc[q] = 0u;
  }
  uint32_t noff, nA, lacc_nyq = 0u, totc_nyq = 0u;
  source_of(NC, noff, nA);

  const int kmin = wv.kmin, kmax = wv.kmax;
  uint2* const stage_t = sc.stage + (size_t)blockIdx.y * wv.rows * NBP;  // this track's rows of the wave
  const int we_rel = (int)min(wv.we - a, (long long)0x3fffffff);  // frames >= a + we_rel lie past the wave end
  const int b_rel = (int)(b - a);

  for (int bi = 0; bi < nbatch; ++bi) {
    const int f_rel = bi * G - 1;                       // first frame of the batch, relative to a
    const long long f_first = a + f_rel;
    mbar_wait(mbar, bi & 1);

    // ---- FFT role: window, radix-16 stage 0 -> padded exchange -> radix-16 stage 1 -> linear store
    if (f_first + g >= 0 && f_rel + g < b_rel) {
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
#if MLX_KA_TAB1
      twiddle_table16(x, tb.tw1_d, t & 15);  // as the general kernel (bit-identical results)
#else
      twiddle_powers<16>(x, tw1.w[0]);
#endif
      dft16<-1>(x);
      const int k = t & 15;
      C* p = fb + (t - k) * 16 + k;  // stage-1 butterfly t writes elements (t-k)*16 + k + 16 r
#pragma unroll
      for (int r = 0; r < 16; ++r) p[r * 16] = x[r];
    }
    __syncthreads();  // sub-transforms of all frames in place; the tile has been consumed
    if (tid == 0 && bi + 1 < nbatch) {
      fence_proxy_async();  // generic-proxy reads of the tile (ordered by the barrier) before the async-proxy refill
      mbar_expect_tx(mbar, TILE * sizeof(float));
      tma_load_1d(tile, tr.x + (f_first + G - 3) * H, TILE * sizeof(float), mbar);
    }

    // ---- pair role: tail butterfly, mirrored partner by shuffle, pair split, analysis, records
    const int g_hi = min(G, b_rel - f_rel);  // frames of this batch that exist: [g_lo, g_hi)
    const int g_lo = f_first < 0 ? 1 : 0;
    //      U frames at a time, stage by stage, so that the independent work of different frames and bins
    //      (loads, butterflies, the sqrt / atan2 of every bin) overlaps; only the cheap phase-advance step at
    //      the end is sequential in the frame index.
    constexpr int U = kKa2Unroll;
    static_assert(G % U == 0, "frames per batch must be a multiple of the bin-phase unroll");
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += U) {
      if (g0 >= g_hi) break;
      C v[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const C* zb = buf + (g0 + u) * BUF + bj;
#pragma unroll
        for (int r = 0; r < R; ++r) v[u][r] = zb[r * NQ];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        twiddle_powers<R>(v[u], wtail);
        dft_r<R, -1>(v[u]);  // v[u][d] = Z[bj + 256 d] of frame g0 + u
      }
      if (tid == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          s_sp[2 * (g0 + u)] = v[u][0];
          s_sp[2 * (g0 + u) + 1] = v[u][R / 2];
        }
      }
      C xk[U][SLOTS], xm[U][SLOTS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          // slot s pairs Z[bj + 256 s] with the other lane's Z[(256 - bj) + 256 (R-1-s)]; the self-mirrored
          // butterflies pair their own outputs: 128 as d <-> R-1-d, 0 as d <-> R-d
          const C mine = (tid == 0) ? v[u][(R - s) % R] : v[u][R - 1 - s];
          const int src_lane = special ? lane : (lane ^ 16);
          C zz;
          zz.x = __shfl_sync(0xffffffffu, mine.x, src_lane);
          zz.y = __shfl_sync(0xffffffffu, mine.y, src_lane);
          const C za = v[u][s];
          // (the general kernel takes bin 256 s of thread 0 from the table, not by rotation: same here)
          const C w = (tid == 0 && s > 0) ? tb.twr_d[NQ * s] : pair_twiddle<N>(wpair0, s);
          const double er = 0.5 * (za.x + zz.x), ei = 0.5 * (za.y - zz.y);
          const double dr = 0.5 * (za.x - zz.x), di = 0.5 * (za.y + zz.y);
          const double tr_ = dr * w.x - di * w.y, ti_ = dr * w.y + di * w.x;
          xk[u][s] = C{er + ti_, ei - tr_};
          xm[u][s] = C{er - ti_, -ei - tr_};
        }
      }
      float mk[U][SLOTS], mm[U][SLOTS];
      uint32_t pk[U][SLOTS], pm[U][SLOTS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          analysis_polar(xk[u][s].x, xk[u][s].y, mk[u][s], pk[u][s]);
          analysis_polar(xm[u][s].x, xm[u][s].y, mm[u][s], pm[u][s]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int gg = g0 + u;
        if (gg >= g_lo && gg < g_hi) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            if (s == 0 && tid == 0) continue;  // bins 0 and NC: left over
            const int k = bj + NQ * s, mbin = NC - k;
            int dk, dm;
            bool fk, fm;
            analysis_advance(xk[u][s].x, xk[u][s].y, sk[s].x.x, sk[s].x.y, pk[u][s], sk[s].p, mk[u][s], sk[s].mag, k,
                             false, dk, fk);
            analysis_advance(xm[u][s].x, xm[u][s].y, sm[s].x.x, sm[s].x.y, pm[u][s], sm[s].p, mm[u][s], sm[s].mag, mbin,
                             false, dm, fm);
            sk[s] = BinState{xk[u][s], pk[u][s], mk[u][s]};
            sm[s] = BinState{xm[u][s], pm[u][s], mm[u][s]};
            // both records into the slot this thread loaded v[s]'s input from (nobody else reads it)
            *reinterpret_cast<uint4*>(buf + gg * BUF + bj + s * NQ) =
                make_uint4(__float_as_uint(mk[u][s]) | (fk ? 0x80000000u : 0u), (uint32_t)dk,
                           __float_as_uint(mm[u][s]) | (fm ? 0x80000000u : 0u), (uint32_t)dm);
          }
        }
      }
    }
    __syncthreads();  // records of the batch complete

    const int e_lo = bi == 0 ? 1 : 0;  // the chunk's leading halo frame emits nothing
    uint2* const row0 = stage_t + (size_t)(f_first - wv.wb) * NBP;  // row of the batch's first frame

    // ---- gather role: bin shift, exact phase increment, chunk-local running phase, coalesced stores
    {
      const int g_cnt = min(g_hi, we_rel - f_rel);  // frames before the wave end
      uint2* const pst = row0 + tid;
#pragma unroll
      for (int gg = 0; gg < G; ++gg) {
        if (gg >= e_lo && gg < g_hi) {
          const unsigned char* fbase = reinterpret_cast<const unsigned char*>(buf + gg * BUF);
#pragma unroll
          for (int q = 0; q < QB; ++q) {
            if (goff[q] != 0xffffffffu) {
              const uint2 rec = *reinterpret_cast<const uint2*>(fbase + goff[q]);
              lacc[q] += shift_inc_a(gA[q], (int)rec.y, rec.x, r_fix);
              if (gg == g_cnt - 1) totc[q] = lacc[q];
              pst[gg * NBP + q * THREADS] = make_uint2(rec.x & 0x7fffffffu, lacc[q]);
            }
          }
        }
      }
    }
    // ---- the Nyquist output bin j = NC: the last warp, one lane per frame, warp scan for the running phase
    if (warp == THREADS / 32 - 1 && noff != 0xffffffffu) {
      const int gg = lane;
      const bool valid = gg < G && gg >= e_lo && gg < g_hi;
      uint32_t inc = 0u, mbits = 0u;
      if (valid) {
        const uint2 rec = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(buf + gg * BUF) + noff);
        inc = shift_inc_a(nA, (int)rec.y, rec.x, r_fix);
        mbits = rec.x & 0x7fffffffu;
      }
      uint32_t run = inc;  // inclusive scan over the lanes (= frames, ascending)
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const uint32_t vv = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += vv;
      }
      const uint32_t mine = lacc_nyq + run;
      if (valid) row0[gg * NBP + NC] = make_uint2(mbits, mine);
      const unsigned cm = __ballot_sync(0xffffffffu, valid && f_rel + gg < we_rel);
      if (cm) totc_nyq = __shfl_sync(0xffffffffu, mine, 31 - __clz(cm));
      lacc_nyq += __shfl_sync(0xffffffffu, run, 7);
    }

    // ---- the three left-over bins on the `spw` warp.  Magnitude and phase of a frame do not depend on its
    //      predecessor, so the frames go in parallel; the phase advance takes the predecessor's values from
    //      the lane below (or the carried state), the running phase is a subgroup scan.
    if (warp == spw) {
      const int sub = lane >> 3, fr = lane & 7;
      const bool valid = sub < 3 && fr >= g_lo && fr < g_hi;
      const bool emitf = valid && fr >= e_lo;
      C X{1.0, 0.0};
      float mag = 1.f;
      uint32_t P = 0u;
      if (valid) {
        if (sub == 0) {  // bin NC/2 mirrors onto itself: X = conj(Z[NC/2]) -- through the pair-split formula
                         // with za = zc and the general kernel's (rotated) twiddle, to stay bit-identical to it
          const C zs = s_sp[2 * fr + 1];
          const double di = 0.5 * (zs.y + zs.y), er = 0.5 * (zs.x + zs.x), ei = 0.5 * (zs.y - zs.y);
          const double dr = 0.5 * (zs.x - zs.x);
          const double tr_ = dr * wself.x - di * wself.y, ti_ = dr * wself.y + di * wself.x;
          X = C{er + ti_, ei - tr_};
          analysis_polar(X.x, X.y, mag, P);
        } else {         // DC and Nyquist are real: X[0] = Re Z0 + Im Z0, X[NC] = Re Z0 - Im Z0
          const C z0 = s_sp[2 * fr];
          X = C{sub == 1 ? z0.x + z0.y : z0.x - z0.y, 0.0};
          mag = fabsf((float)X.x);
          P = X.x < 0.0 ? 0x80000000u : 0u;
        }
      }
      C Xp;
      Xp.x = __shfl_up_sync(0xffffffffu, X.x, 1, 8);
      Xp.y = __shfl_up_sync(0xffffffffu, X.y, 1, 8);
      uint32_t Pp = __shfl_up_sync(0xffffffffu, P, 1, 8);
      float mp = __shfl_up_sync(0xffffffffu, mag, 1, 8);
      if (fr == g_lo) {
        Xp = ssp.x;
        Pp = ssp.p;
        mp = ssp.mag;
      }
      int dq = 0;
      bool flip = false;
      if (sub == 0) {
        analysis_advance(X.x, X.y, Xp.x, Xp.y, P, Pp, mag, mp, NC / 2, false, dq, flip);
      } else {  // analysis_real_bin: arg X is 0 or pi, Im Z := +0, so d is 0 or +pi (stored as -2^31 with the flip flag)
        flip = (P != Pp) && !(mag * mp <= 1e-18f);
        dq = flip ? (int)0x80000000u : 0;
      }
      uint32_t run = 0u;
      if (emitf) run = shift_inc_a(spA, dq, __float_as_uint(mag) | (flip ? 0x80000000u : 0u), r_fix);
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const uint32_t vr = __shfl_up_sync(0xffffffffu, run, o, 8);
        if (fr >= o) run += vr;
      }
      const uint32_t lacc_f = sp_lacc + run;
      if (emitf && sp_fed) {
        row0[fr * NBP + spj] = make_uint2(__float_as_uint(mag), lacc_f);
        if (f_rel + fr == we_rel - 1) sc.totc[trow + spj] = lacc_f;  // phase the next wave starts from
      }
      if (g_hi > g_lo) {  // carry = the batch's last frame
        const int last = (lane & 24) + g_hi - 1;
        ssp.x.x = __shfl_sync(0xffffffffu, X.x, last);
        ssp.x.y = __shfl_sync(0xffffffffu, X.y, last);
        ssp.p = __shfl_sync(0xffffffffu, P, last);
        ssp.mag = __shfl_sync(0xffffffffu, mag, last);
        sp_lacc = __shfl_sync(0xffffffffu, lacc_f, last);
      }
    }

    // ---- peak bin (lowest k on exact ties) and f0, one warp per frame, from the records of the band
    for (int gg = warp + e_lo; gg < G; gg += THREADS / 32) {
      if (f_rel + gg >= b_rel || f_rel + gg >= we_rel) break;
      const long long ff = f_first + gg;
      const uint2* recs = reinterpret_cast<const uint2*>(buf + gg * BUF);  // record of bin k < NC/2: recs[2 k]
      float best = -1.f;
      int bk = kmin;
      for (int k = kmin + lane; k <= kmax; k += 32) {
        const float vv = __uint_as_float(recs[2 * k].x & 0x7fffffffu);
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
          const uint2 rb = recs[2 * bk];
          long long dd = (long long)(int)rb.y;
          if (rb.x >> 31) dd += (dd < 0) ? 4294967296LL : -4294967296LL;
          tr.f0[ff] = ((float)bk + (float)dd * 9.313225746154785e-10f) * wv.fs_over_N;  // nu = k + 4 d
        }
      }
    }
    __syncthreads();  // the frame buffers are rewritten by the next batch
  }

  // ---- chunk totals: all frames (prefix of the later chunks of this wave) and the frames before the wave end
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    if (goff[q] != 0xffffffffu) {
      sc.tot[trow + tid + q * THREADS] = lacc[q];
      sc.totc[trow + tid + q * THREADS] = totc[q];
    }
  }
  if (tid == THREADS - 1 && noff != 0xffffffffu) {
    sc.tot[trow + NC] = lacc_nyq;
    sc.totc[trow + NC] = totc_nyq;
  }
  if (sp_owner && sp_fed) {
    sc.tot[trow + spj] = sp_lacc;
    if (b <= wv.we) sc.totc[trow + spj] = sp_lacc;   // every frame of the chunk lies before the wave end
    else if (a >= wv.we) sc.totc[trow + spj] = 0u;    // none does (otherwise: captured at frame we - 1)
  }
}

// ------------------------------------------------------------------------------------------------
bool pv_analyze2_supported(int fftN) { return fftN == 1024 || fftN == 2048; }
int pv_analyze2_band_capacity(int fftN) { return fftN / 8 - 1; }  // the band must lie below bin fftN/8

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
