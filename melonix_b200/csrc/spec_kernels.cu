// melonix_b200/csrc/spec_kernels.cu -- K1: batched STFT-magnitude frames for sm_100a.
//
// Replaces Spec::internalGetSpec (reference spec.cpp:44-66): per job (start, end) take the window
// [end-N, end), zero outside [0, n), multiply samples before `start` by expf(-2.5e-4f*(start-i))
// (float product, spec.cpp:58), transform, and emit |X[k]|/N for k in [0, N/2) (spec.cpp:61-65).
// The reference runs one 32768-point complex double FFTW transform per job on one thread; here a
// group of N/32 threads runs the real transform as an N/2-point complex FP32 Stockham FFT in shared
// memory (fft.cuh) and many jobs are in flight per launch.  K7 (optional) fuses the colour ramp of
// SpecCache::populateTex (reference spec-cache.cpp:77-96) into the epilogue.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "fft.cuh"
#include "kernels.h"
#include "spec_frame.cuh"
#include "tma.cuh"

namespace mlx {

template <int N>
struct SpecCfg {
  static constexpr int NC = N / 2;
  static constexpr int TPF = NC / 16;
  static constexpr int JPB = TPF >= 256 ? 1 : 256 / TPF;  // jobs in flight per CTA
  static constexpr int THREADS = JPB * TPF;
  static constexpr int BUF = FftPlan<NC>::BUF;
  static constexpr bool TAB = (N <= 8192);  // window-factor table staged in shared memory
  static constexpr size_t SMEM = sizeof(cplx<float>) * JPB * BUF + (TAB ? sizeof(float) * (N + 4) : 0);
};

template <int TPF>
struct SpecBar {
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

// colour ramp, reference spec-cache.cpp:79-96 (note 255/3 == 85 and 2*255/3 == 170: int division)
__device__ __forceinline__ void colour_ramp(float v, float k, unsigned char* rgb) {
  float tmp = v * k;
  tmp = tmp < 0.f ? 0.f : (255.f < tmp ? 255.f : tmp);
  unsigned char r, g, b;
  if (tmp < 85.f) {
    r = (unsigned char)__float2int_rz(tmp);
    g = 0;
    b = 0;
  } else if (tmp < 170.f) {
    const double a = (double)((tmp - 85.f) / 85.f) * 3.141592 / 2;
    r = (unsigned char)__double2int_rz((double)tmp * cos(a));
    g = (unsigned char)__double2int_rz((double)tmp * sin(a));
    b = 0;
  } else {
    const unsigned char lk = (unsigned char)__float2int_rz((tmp - 170.f) * 3.f);
    r = lk;
    g = (unsigned char)__float2int_rz(tmp);
    b = lk;
  }
  rgb[0] = r;
  rgb[1] = g;
  rgb[2] = b;
}

#ifndef MLX_SPEC_PRE1
#define MLX_SPEC_PRE1 0  // K1r: stage-1 twiddle powers in registers (needs MLX_SPEC_MINB=2 to avoid spills)
#endif
#ifndef MLX_SPEC_MINB
#define MLX_SPEC_MINB 3
#endif

template <int N>
__global__ void __launch_bounds__(SpecCfg<N>::THREADS, SpecCfg<N>::THREADS >= 512 ? 1 : MLX_SPEC_MINB)
spec_kernel(const SpecArgs a) {
  using Cfg = SpecCfg<N>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, JPB = Cfg::JPB, BUF = Cfg::BUF;
  using C = cplx<float>;
  using F = Fft<float, NC, -1>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* bufs = reinterpret_cast<C*>(smem_raw);
  float* s_decay = reinterpret_cast<float*>(bufs + JPB * BUF);  // [N + 1] when Cfg::TAB

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  C* buf = bufs + g * BUF;
  if constexpr (Cfg::TAB) {
    for (int d = tid; d <= N; d += Cfg::THREADS) s_decay[d] = a.decay[d];
    __syncthreads();
  }
  FftTwiddles<float, NC, -1> twd;
  twd.init(t, a.tw_f);
  unsigned mask = 0xffffffffu;
  if constexpr (TPF < 32) mask = ((1u << TPF) - 1u) << (((tid & 31) / TPF) * TPF);
  const SpecBar<TPF> bar{1 + g, mask};
  const float inv_n = 1.0f / (float)N;

  for (long long job = (long long)blockIdx.x * JPB + g; job < a.count; job += (long long)gridDim.x * JPB) {
    long long start, end;
    if (a.jobs) {
      start = a.jobs[2 * job];
      end = a.jobs[2 * job + 1];
    } else {
      start = (a.first_frame + job) * a.hop;
      end = start + a.hop;
    }
    C x[16];
    // window [end-N, end): zero outside [0, n); samples before `start` are multiplied by
    // expf(-2.5e-4f * (start - i)) -- a float product, as spec.cpp:58.  The factor depends only on
    // the integer distance start - i <= N; for N <= 8192 it comes from a shared-memory copy of a
    // table the host filled with glibc's expf (the reference's own arithmetic).
    const long long base = end - N;
    const bool interior = base >= 0 && end <= a.n;  // whole window inside the track: no bounds tests
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int p = 2 * (t + m * TPF);  // position inside the window
      float v[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const long long i = base + p + c;
        float s = 0.f;
        if (interior || (i >= 0 && i < a.n)) s = __ldg(a.x + i);
        const long long dist = start - i;  // > 0: before `start`
        if (dist > 0) {
          float w;
          if (Cfg::TAB && dist <= N) w = s_decay[dist];
          else if (dist <= N) w = __ldg(a.decay + dist);  // fftN > 8192: the same host table, through L1/L2
          else w = expf(-2.5e-4f * (float)(int)dist);      // only for jobs with end < start
          s = __fmul_rn(w, s);
        }
        v[c] = s;
      }
      x[m] = C{v[0], v[1]};
    }
    F::run(x, buf, t, twd, bar);
    F::store(x, buf, t);
    bar.sync();
    float* out = a.out ? a.out + job * NC : nullptr;
    unsigned char* rgb = a.rgb ? a.rgb + job * NC * 3 : nullptr;
    // bins k and NC-k from Z[k], Z[NC-k]; k = t + q*TPF in [0, NC/2]
#pragma unroll 1
    for (int q = 0; q < 9; ++q) {
      const int k = t + q * TPF;
      if (k > NC / 2) break;
      const C za = buf[fft_pad(k)];
      const C zc = buf[fft_pad((NC - k) & (NC - 1))];
      float mk, mm;
      if (k == 0) {
        mk = fabsf(za.x + za.y);  // X[0]; the Nyquist bin X[NC] is dropped (spec.cpp:61)
        mm = 0.f;
      } else {
        const C w = a.twr_f[k];
        const float er = 0.5f * (za.x + zc.x), ei = 0.5f * (za.y - zc.y);
        const float dr = 0.5f * (za.x - zc.x), di = 0.5f * (za.y + zc.y);
        const float tr_ = dr * w.x - di * w.y, ti_ = dr * w.y + di * w.x;
        const float xkr = er + ti_, xki = ei - tr_;
        const float xmr = er - ti_, xmi = -ei - tr_;
        mk = spec_sqrt(fmaf(xkr, xkr, xki * xki));
        mm = spec_sqrt(fmaf(xmr, xmr, xmi * xmi));
      }
      mk *= inv_n;
      mm *= inv_n;
      if (out) {
        out[k] = mk;
        if (k != 0 && k != NC - k) out[NC - k] = mm;
      }
      if (rgb) {
        colour_ramp(mk, a.kcol, rgb + 3 * k);
        if (k != 0 && k != NC - k) colour_ramp(mm, a.kcol, rgb + 3 * (NC - k));
      }
    }
    bar.sync();  // buf is rewritten by the next job's first stage
  }
}

// ------------------------------------------------------------------------------------------------
// K1r: the regular-hop form of K1 (job f = (start = f*hop, end = start + hop), the geometry
// SpecCache::populateTex produces at a fixed zoom, spec-cache.cpp:63-65).  Consecutive frames share
// N - hop samples, so a CTA takes a run of consecutive frames and stages one contiguous sample tile
// per batch of JPB frames with a 1-D TMA bulk copy (double-buffered: the copy of batch b+2 is in
// flight while batch b is transformed).  The window factor depends only on the position inside the
// frame (distance to `start` = N - hop - p), so it is one shared-memory table and one float product
// per sample -- exactly the product of spec.cpp:58 -- and the zero padding kept around every track
// in HBM replaces the bounds tests of spec.cpp:50-54.  The last FFT stage leaves Z in registers;
// only the upper half of the slots goes through shared memory to reach the thread that owns the
// mirrored bin.
template <int N>
struct SpecFramesCfg {
  static constexpr int NC = N / 2;
  static constexpr int TPF = NC / 16;
  static constexpr int THREADS = 256;
  static constexpr int JPB = THREADS / TPF;
  static constexpr int BUF = FftPlan<NC>::BUF;
  static constexpr bool TAB = (N <= 2048);  // window + split-twiddle tables staged in shared memory
  static constexpr int TWR = ((NC / 2 + 1) + 1) & ~1;
  static_assert(TPF <= THREADS, "regular-hop Spec kernel: fftN <= 8192");
  static size_t smem(int hop) {
    return sizeof(cplx<float>) * JPB * BUF + (TAB ? sizeof(float) * N + sizeof(cplx<float>) * TWR : 0) +
           sizeof(float) * 2 * (size_t)(N + (JPB - 1) * hop) + 16;
  }
};

template <int N, bool RGB>
__global__ void __launch_bounds__(256, MLX_SPEC_MINB)
spec_frames_kernel(const SpecArgs a0, const int fpc) {
  SpecArgs a = a0;
  if (a.multi) {  // batched launch: this CTA row works on track blockIdx.y
    const SpecTrackDesc d = a.multi[blockIdx.y];
    a.x = d.x;
    a.n = d.n;
    a.count = d.count;
    a.out = d.out;
    a.rgb = d.rgb;
  }
  using Cfg = SpecFramesCfg<N>;
  constexpr int NC = Cfg::NC, TPF = Cfg::TPF, JPB = Cfg::JPB, BUF = Cfg::BUF, THREADS = Cfg::THREADS;
  constexpr bool TAB = Cfg::TAB;
  using C = cplx<float>;
  using SF = SpecFrame<N>;
  using F = typename SF::F;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* bufs = reinterpret_cast<C*>(smem_raw);                               // [JPB][BUF]
  float* s_win = reinterpret_cast<float*>(bufs + JPB * BUF);              // [N]         (TAB)
  C* s_twr = reinterpret_cast<C*>(s_win + (TAB ? N : 0));                 // [TWR]       (TAB)
  float* tile = reinterpret_cast<float*>(s_twr + (TAB ? Cfg::TWR : 0));   // [2][span]
  const int hop = a.hop;
  const int span = N + (JPB - 1) * hop;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tile + 2 * span);          // [2]

  const int tid = threadIdx.x;
  const int g = tid / TPF, t = tid % TPF;
  const long long f_begin = (long long)blockIdx.x * fpc;
  if (f_begin >= a.count) return;
  const int nfr = (int)min((long long)fpc, a.count - f_begin);
  const int nbatch = (nfr + JPB - 1) / JPB;
  // first sample of the tile of batch b: window start of its first frame, (f + 1) * hop - N
  const float* src0 = a.x + (a.first_frame + f_begin + 1) * hop - N;
  const uint32_t tile_bytes = (uint32_t)span * sizeof(float);

  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
  }
  const int ndec = N - hop;  // samples of the frame that lie before `start`
  if constexpr (TAB && !MLX_SPEC_DERIVE) {
    for (int p = tid; p < N; p += THREADS) s_win[p] = p < ndec ? a.decay[ndec - p] : 1.f;
    for (int k = tid; k <= NC / 2; k += THREADS) s_twr[k] = a.twr_f[k];
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(mbar, tile_bytes);
    tma_load_1d(tile, src0, tile_bytes, mbar);
    if (nbatch > 1) {
      mbar_expect_tx(mbar + 1, tile_bytes);
      tma_load_1d(tile + span, src0 + (long long)JPB * hop, tile_bytes, mbar + 1);
    }
  }

  C* buf = bufs + g * BUF;
  FftTwiddles<float, NC, -1, MLX_SPEC_PRE1 != 0> twd;
  twd.init(t, a.tw_f);
  unsigned mask = 0xffffffffu;
  if constexpr (TPF < 32) mask = ((1u << TPF) - 1u) << (((tid & 31) / TPF) * TPF);
  const SpecBar<TPF> bar{1 + g, mask};
  const float scale = 0.5f / (float)N;  // the pair split below works on 2 X
#if MLX_SPEC_DERIVE
  // the two exact window factors and the one split twiddle everything else of this thread is derived from
  const float wc0 = 2 * t < ndec ? __ldg(a.decay + (ndec - 2 * t)) : 1.f;
  const float wc1 = 2 * t + 1 < ndec ? __ldg(a.decay + (ndec - 2 * t - 1)) : 1.f;
  const float2 w0v = __ldg(reinterpret_cast<const float2*>(a.twr_f + t));
  const C w0{w0v.x, w0v.y};
#endif

  for (int b = 0; b < nbatch; ++b) {
    const int fr = b * JPB + g;
    const bool active = fr < nfr;
    mbar_wait(mbar + (b & 1), (b >> 1) & 1);
    C x[16];
    if (active) {
      const float* cur = tile + (b & 1) * span + g * hop;
#if MLX_SPEC_DERIVE
      if constexpr (true) {
        SF::load_derived(x, cur, t, wc0, wc1);
      } else
#endif
      if constexpr (TAB) {
        SF::load(x, cur, t, [&](int p) { return *reinterpret_cast<const C*>(s_win + p); });
      } else {
        SF::load(x, cur, t, [&](int p) {
          return C{p < ndec ? __ldg(a.decay + (ndec - p)) : 1.f, p + 1 < ndec ? __ldg(a.decay + (ndec - p - 1)) : 1.f};
        });
      }
    }
    __syncthreads();  // every group has taken its samples out of tile (b & 1); previous epilogue reads done
    if (tid == 0 && b + 2 < nbatch) {
      fence_proxy_async();  // generic-proxy reads of the tile (ordered by the barrier) before the async-proxy refill
      mbar_expect_tx(mbar + (b & 1), tile_bytes);
      tma_load_1d(tile + (b & 1) * span, src0 + (long long)(b + 2) * JPB * hop, tile_bytes, mbar + (b & 1));
    }
    if (!active) continue;
    F::run(x, buf, t, twd, bar);
    bar.sync();  // every thread of the group has loaded its last-stage inputs: the staging area may be overwritten
    SF::stage_upper(x, buf, t);
    bar.sync();
    const long long frame = f_begin + fr;
    float* out = a.out ? a.out + frame * NC : nullptr;
    unsigned char* rgb = RGB ? a.rgb + frame * NC * 3 : nullptr;
    auto twr = [&](int k) {
      if constexpr (TAB) {
        return s_twr[k];
      } else {
        const float2 wv = __ldg(reinterpret_cast<const float2*>(a.twr_f + k));
        return C{wv.x, wv.y};
      }
    };
    auto emit = [&](int k, float v) {
      if (out) out[k] = v;
      if constexpr (RGB) colour_ramp(v, a.kcol, rgb + 3 * k);
    };
#if MLX_SPEC_DERIVE
    SF::template emit_bins<true>(x, buf, t, scale, twr, emit, w0);
#else
    SF::emit_bins(x, buf, t, scale, twr, emit);
#endif
  }
}

#define MLX_SPEC_DISPATCH(N_, ...)                           \
  switch (N_) {                                              \
    case 512: { constexpr int N = 512; __VA_ARGS__; } break;     \
    case 1024: { constexpr int N = 1024; __VA_ARGS__; } break;   \
    case 2048: { constexpr int N = 2048; __VA_ARGS__; } break;   \
    case 4096: { constexpr int N = 4096; __VA_ARGS__; } break;   \
    case 8192: { constexpr int N = 8192; __VA_ARGS__; } break;   \
    case 16384: { constexpr int N = 16384; __VA_ARGS__; } break; \
    case 32768: { constexpr int N = 32768; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;                   \
  }

#define MLX_SPECR_DISPATCH(N_, ...)                          \
  switch (N_) {                                              \
    case 512: { constexpr int N = 512; __VA_ARGS__; } break;     \
    case 1024: { constexpr int N = 1024; __VA_ARGS__; } break;   \
    case 2048: { constexpr int N = 2048; __VA_ARGS__; } break;   \
    case 4096: { constexpr int N = 4096; __VA_ARGS__; } break;   \
    case 8192: { constexpr int N = 8192; __VA_ARGS__; } break;   \
    default: return cudaErrorInvalidValue;                   \
  }

constexpr size_t kMaxDynSmem = 227 * 1024;

template <int N, bool RGB>
static cudaError_t configure_frames() {
  size_t m = SpecFramesCfg<N>::smem(N);  // hop <= N
  if (m > kMaxDynSmem) m = kMaxDynSmem;
  return cudaFuncSetAttribute(spec_frames_kernel<N, RGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m);
}

cudaError_t spec_configure(int fftN) {
  MLX_SPEC_DISPATCH(fftN, {
    cudaError_t e = cudaFuncSetAttribute(spec_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SpecCfg<N>::SMEM);
    if (e != cudaSuccess) return e;
  });
  if (fftN <= 8192) {
    MLX_SPECR_DISPATCH(fftN, {
      cudaError_t e = configure_frames<N, false>();
      if (e != cudaSuccess) return e;
      return configure_frames<N, true>();
    });
  }
  return cudaSuccess;
}

// Regular-hop launches go to K1r when its tiles are aligned for TMA and stay inside the zero padding
// of the track; everything else (arbitrary job lists, odd hops such as the reference's 375-sample
// pixel, fftN > 8192) takes the general kernel.
static bool spec_frames_eligible(int fftN, const SpecArgs& a) {
  static const bool off = getenv("MLX_SPEC_GENERIC") != nullptr;
  if (off || a.jobs != nullptr || fftN > 8192) return false;
  if (a.hop <= 0 || a.hop > fftN || (a.hop & 3) != 0 || a.first_frame < 0) return false;
  const int jpb = 256 / (fftN / 32);
  if ((a.first_frame + a.count + jpb) * (long long)a.hop > a.n + kPadBack) return false;
  return true;
}

template <int N>
static cudaError_t launch_spec_frames(const SpecArgs& a, cudaStream_t st) {
  using Cfg = SpecFramesCfg<N>;
  const size_t smem = Cfg::smem(a.hop);
  if (smem > kMaxDynSmem) return cudaErrorInvalidConfiguration;
  int dev = 0, sms = 148, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (a.rgb) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spec_frames_kernel<N, true>, Cfg::THREADS, smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spec_frames_kernel<N, false>, Cfg::THREADS, smem);
  if (occ < 1) occ = 1;
  // frames per CTA: whole batches, at most 32 of them, and a grid that fills the resident slots evenly
  // (batched launch: a.count is the longest track's frame count; the CTA rows of all tracks share the slots)
  const long long batches = (a.count + Cfg::JPB - 1) / Cfg::JPB * (a.multi ? a.ntracks : 1);
  const long long slots = (long long)sms * occ;
  const long long rounds = (batches + slots * 32 - 1) / (slots * 32);
  long long nb = (batches + slots * rounds - 1) / (slots * rounds);
  if (nb < 1) nb = 1;
  const int fpc = (int)nb * Cfg::JPB;
  const dim3 grid((unsigned)((a.count + fpc - 1) / fpc), a.multi ? (unsigned)a.ntracks : 1u);
  if (a.rgb) spec_frames_kernel<N, true><<<grid, Cfg::THREADS, smem, st>>>(a, fpc);
  else spec_frames_kernel<N, false><<<grid, Cfg::THREADS, smem, st>>>(a, fpc);
  return cudaGetLastError();
}

cudaError_t launch_spec(int fftN, const SpecArgs& a, cudaStream_t st) {
  if (a.count <= 0) return cudaSuccess;
  if (spec_frames_eligible(fftN, a)) {
    cudaError_t e = cudaErrorInvalidConfiguration;
    MLX_SPECR_DISPATCH(fftN, e = launch_spec_frames<N>(a, st));
    if (e != cudaErrorInvalidConfiguration) return e;
    (void)cudaGetLastError();
  }
  MLX_SPEC_DISPATCH(fftN, {
    using Cfg = SpecCfg<N>;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spec_kernel<N>, Cfg::THREADS, Cfg::SMEM);
    if (occ < 1) occ = 1;
    long long want = (a.count + Cfg::JPB - 1) / Cfg::JPB;
    long long cap = (long long)sms * occ * 4;  // a few CTAs per resident slot, grid-stride over jobs
    const int grid = (int)(want < cap ? want : cap);
    spec_kernel<N><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
  });
  return cudaGetLastError();
}

}  // namespace mlx
