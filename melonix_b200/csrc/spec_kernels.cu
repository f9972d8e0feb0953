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

#include "fft.cuh"
#include "kernels.h"

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

// sqrt.approx: ~1 ulp, far inside the 1e-4 RMS budget of the float magnitudes
__device__ __forceinline__ float spec_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
          else w = expf(-2.5e-4f * (float)(int)dist);
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

cudaError_t spec_configure(int fftN) {
  MLX_SPEC_DISPATCH(fftN, return cudaFuncSetAttribute(spec_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int)SpecCfg<N>::SMEM));
  return cudaSuccess;
}

cudaError_t launch_spec(int fftN, const SpecArgs& a, cudaStream_t st) {
  if (a.count <= 0) return cudaSuccess;
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
