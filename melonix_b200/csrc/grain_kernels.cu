// melonix_b200/csrc/grain_kernels.cu -- K6: grain resampler + float->int16 for sm_100a.
//
// Replaces the per-sample loop of App::process (reference app.cpp:331-343) and the conversion loop
// of App::exportWav (app.cpp:1209-1212).  The reference walks one grain at a time, pushing
//   (1 - frac) * grain[idx] + frac * (idx + 1 < size ? grain[idx + 1] : nextGrainFirstSample)
// with idx/frac = modf(float(i) * rate + bias) (bias is always 0.f, app.hpp:66), all in float.
// The schedule (which grain, which rate, where in the output) is a short serial recurrence over
// grains that stays on the host (melonix_b200/host/grain_schedule.hpp); given it, every output
// sample is independent: one thread per sample, binary search of the schedule, two gathered loads.
// Float products and the final sum are kept un-fused (__fmul_rn/__fadd_rn) so the result is
// bit-identical to the reference's scalar x86 arithmetic.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mlx {

__global__ void __launch_bounds__(256) grain_kernel(const GrainArgs a) {
  const long long rendered = a.out_off[a.ngrains];
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < a.total;
       o += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;  // tail: preferredGrainSize zeros pushed when no grain is left (app.cpp:303-309)
    if (o < rendered) {
      int lo = 0, hi = a.ngrains - 1;  // largest g with out_off[g] <= o
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.out_off[mid] <= o) lo = mid; else hi = mid - 1;
      }
      const int i = (int)(o - a.out_off[lo]);
      const float p = __fadd_rn(__fmul_rn((float)i, a.g_rate[lo]), 0.f);
      const float idxF = truncf(p);
      const float frac = __fsub_rn(p, idxF);
      const int idx = (int)idxF;
      const float* gx = a.x + a.g_start[lo];
      const float x0 = gx[idx];
      const float x1 = (idx + 1 < a.g_len[lo]) ? gx[idx + 1] : a.g_next[lo];
      v = __fadd_rn(__fmul_rn(__fsub_rn(1.f, frac), x0), __fmul_rn(frac, x1));
    }
    if (a.out) a.out[o] = v;
    if (a.out_i16) a.out_i16[o] = (short)__double2int_rz((double)v * 32767.);
  }
}

cudaError_t launch_grain(const GrainArgs& a, cudaStream_t st) {
  if (a.total <= 0) return cudaSuccess;
  long long blocks = (a.total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  grain_kernel<<<(int)blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace mlx
