// melonix_b200/csrc/pv_common.cuh -- compile-time geometry and group barriers shared by the phase-vocoder
// kernels (pv_kernels.cu: general analysis K_A, scan, synthesis K_S; pv_analyze2.cu: the constant-rate
// pitch-up analysis kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft.cuh"
#include "kernels.h"

#ifndef MLX_KA_CTAS_MAXN
#define MLX_KA_CTAS_MAXN 2048  // largest fftN analysed with MLX_KA_CTAS CTAs per SM (beyond: one 512-thread CTA)
#endif
#ifndef MLX_KA_CTAS
#define MLX_KA_CTAS 2  // analysis CTAs per SM for small fftN: 2 x 256 threads (measured 5 % faster than 1 x 512:
                       // the FP64 FFT phase of one CTA overlaps the integer/FP32 phases of the other)
#endif

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
  static constexpr int QP = (NC / 2 + THREADS - 1) / THREADS;      // pair slots per thread (k = 1..NC/2)
  static constexpr int QB = (NC + THREADS - 1) / THREADS;          // bin slots per thread (bins 0..NC-1; bin NC: last warp)
  static constexpr bool WIN_D = (N <= 2048);                       // double window staged in smem
  static constexpr int BUFS = BUF + 2;  // + the Nyquist bin's (mag, d) record + an all-zero record (empty K_j)
  static constexpr size_t SMEM_A = sizeof(cplx<double>) * G * BUFS + (WIN_D ? sizeof(double) * N : 0) +
                                   sizeof(float) * 2 * TILE + 64;
  static constexpr size_t SMEM_S = sizeof(cplx<float>) * G * BUF + sizeof(uint2) * G * NBP + 16 + 64;  // + stage records
};

// Frames per batch: synthesis keeps G*(N/2) = 4096 complex points in flight (256 threads, 2 CTAs per
// SM); analysis the same for fftN <= MLX_KA_CTAS_MAXN and 8192 points in one 512-thread CTA per SM
// beyond (the per-thread bin state of a 256-thread CTA spills there): 16 resident warps per SM.
template <int N>
struct PvG {
  static constexpr int value = 8192 / N;    // K_S
  static constexpr int ka_ctas = (N <= MLX_KA_CTAS_MAXN) ? MLX_KA_CTAS : 1;  // (two 8192-point CTAs do not fit one SM)
  static constexpr int analyze = (16384 / ka_ctas) / N;           // K_A
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

}  // namespace mlx
