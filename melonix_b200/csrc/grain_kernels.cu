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

#include "grain_seg.cuh"
#include "kernels.h"

namespace mlx {

constexpr int kGrainSegment = 256 * 16;  // output samples per CTA step

__global__ void __launch_bounds__(256) grain_kernel(const GrainArgs a) {
  __shared__ int s_g0;
  const long long rendered = a.out_off[a.ngrains];
  // A CTA renders contiguous segments of the output.  The schedule row of the segment's first sample
  // is found once (binary search, thread 0); a segment spans only a few rows (a row is ~1500 / rate
  // samples), so every thread then walks forward from it instead of searching per sample.
  for (long long seg = (long long)blockIdx.x * kGrainSegment; seg < a.total; seg += (long long)gridDim.x * kGrainSegment) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = a.ngrains - 1;  // largest g with out_off[g] <= seg (rows are non-empty in output time)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.out_off[mid] <= seg) lo = mid; else hi = mid - 1;
      }
      s_g0 = lo;
    }
    __syncthreads();
    int g = s_g0;
    long long next_off = a.ngrains > 0 ? a.out_off[g + 1] : 0;
#pragma unroll 4
    for (int j = 0; j < kGrainSegment / 256; ++j) {
      const long long o = seg + threadIdx.x + 256 * j;
      if (o >= a.total) break;
      float v = 0.f;  // tail: preferredGrainSize zeros pushed when no grain is left (app.cpp:303-309)
      if (o < rendered) {
        while (o >= next_off) {  // rows with zero output samples are skipped like the binary search did
          ++g;
          next_off = a.out_off[g + 1];
        }
        const int i = (int)(o - a.out_off[g]);
        const float p = __fadd_rn(__fmul_rn((float)i, a.g_rate[g]), 0.f);
        const float idxF = truncf(p);
        const float frac = __fsub_rn(p, idxF);
        const int idx = (int)idxF;
        const float* gx = a.x + a.g_start[g];
        const float x0 = gx[idx];
        const float x1 = (idx + 1 < a.g_len[g]) ? gx[idx + 1] : a.g_next[g];
        v = __fadd_rn(__fmul_rn(__fsub_rn(1.f, frac), x0), __fmul_rn(frac, x1));
      }
      if (a.out) a.out[o] = v;
      if (a.out_i16) a.out_i16[o] = (short)__double2int_rz((double)v * 32767.);
    }
    __syncthreads();  // s_g0 is rewritten by the next segment
  }
}

cudaError_t launch_grain(const GrainArgs& a, cudaStream_t st) {
  if (a.total <= 0) return cudaSuccess;
  long long blocks = (a.total + kGrainSegment - 1) / kGrainSegment;
  if (blocks > 148 * 32) blocks = 148 * 32;
  grain_kernel<<<(int)blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K8: grain segmentation (replaces the zero-crossing search of App::preproc, reference
// app.cpp:156-235; bit logic and its derivation in grain_seg.cuh).
//
// K8a  grain_cross_kernel: the two crossing predicates for every sample of every track, data
//      parallel.  A CTA covers 256 words (8192 samples): warps turn samples into sign words (one
//      coalesced float4 load per lane, four sign bits per lane OR-reduced over octets of lanes), the
//      words meet in shared memory and every thread combines four of them into one word of Z7 and one
//      of Z3.  HBM-bound: 4 bytes read and 1/4 byte written per sample.
// K8b  grain_chain_kernel: the chain over grains is inherently serial (every grain starts where the
//      previous one ended) but O(#grains): one CTA per track stages 32 KB slices of Z7 in shared
//      memory (256 Ki samples, ~170 grains per refill) and warp 0 walks it -- 24 lanes x 64 bits
//      around start + 1500, two ballots find the nearest crossing on either side of the centre --
//      falling back to a forward scan of Z3 in global memory when the window holds no crossing
//      (app.cpp:194-231).
constexpr int kSegWordsPerCta = 256;
constexpr int kSegStageWords = 8192;

__global__ void __launch_bounds__(256) grain_cross_kernel(const GrainSegTrack* __restrict__ tracks) {
  // sign words k0-4 .. k0+259 (the CTA's 256 words plus one 128-sample group on each side)
  __shared__ uint32_t sL[kSegWordsPerCta + 8], sR[kSegWordsPerCta + 8];
  const GrainSegTrack tr = tracks[blockIdx.y];
  const long long k0 = (long long)blockIdx.x * kSegWordsPerCta;  // first word of this CTA
  if (k0 >= tr.nwords) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // One warp iteration = 128 samples as one float4 per lane (a coalesced 512-byte load); the lane's
  // four sign bits are shifted to their place inside the word of its octet and OR-reduced over the
  // octet (three butterfly shuffles).  Samples outside [0, n) are read from the zero padding kept around every track and
  // never matter: the range test of seg_cross_word masks every position whose look-around leaves the
  // track.
  for (int grp = warp; grp < kSegWordsPerCta / 4 + 2; grp += 8) {
    const long long i = (k0 - 4 + 4 * grp) * 32 + 4 * lane;
    const float4 v = __ldg(reinterpret_cast<const float4*>(tr.x + i));
    const uint32_t l = (!(v.x >= 0.f) ? 1u : 0u) | (!(v.y >= 0.f) ? 2u : 0u) | (!(v.z >= 0.f) ? 4u : 0u) |
                       (!(v.w >= 0.f) ? 8u : 0u);
    const uint32_t r = (!(v.x < 0.f) ? 1u : 0u) | (!(v.y < 0.f) ? 2u : 0u) | (!(v.z < 0.f) ? 4u : 0u) |
                       (!(v.w < 0.f) ? 8u : 0u);
    const int sh = 4 * (lane & 7);
    uint32_t lw = l << sh, rw = r << sh;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {  // OR over the octet of lanes that shares a word
      lw |= __shfl_xor_sync(0xffffffffu, lw, o);
      rw |= __shfl_xor_sync(0xffffffffu, rw, o);
    }
    if ((lane & 7) == 0) {
      sL[4 * grp + (lane >> 3)] = lw;
      sR[4 * grp + (lane >> 3)] = rw;
    }
  }
  __syncthreads();
  const long long k = k0 + threadIdx.x;
  if (k < tr.nwords) {
    const int w = threadIdx.x + 4;
    tr.z7[k] = seg_cross_word(sL[w - 1], sL[w], sR[w], sR[w + 1], k, tr.n, 7);
    tr.z3[k] = seg_cross_word(sL[w - 1], sL[w], sR[w], sR[w + 1], k, tr.n, 3);
  }
}

__global__ void __launch_bounds__(256) grain_chain_kernel(const GrainSegTrack* __restrict__ tracks, int cap) {
  __shared__ uint32_t stage[kSegStageWords];
  __shared__ int s_start, s_count, s_done;
  const GrainSegTrack tr = tracks[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31;
  const int lim = (int)(tr.n - kGrainPreferred - 1);  // app.cpp:161
  if (tid == 0) {
    s_start = 0;
    s_count = 0;
    s_done = !(0 < lim);
  }
  __syncthreads();
  while (!s_done) {
    int start = s_start;
    // refill: the staged slice begins at the word that holds the first probe of the next grain
    const long long wbase = ((long long)start + kGrainPreferred - kGrainHalfSpan) >> 5;
    for (int i = tid; i < kSegStageWords; i += 256) stage[i] = (wbase + i < tr.nwords) ? __ldg(tr.z7 + wbase + i) : 0u;
    __syncthreads();
    if (tid < 32) {
      int count = s_count, done = 0;
      while (true) {
        if (!(start < lim)) {
          done = 1;
          break;
        }
        const int lo = start + kGrainPreferred - kGrainHalfSpan;  // first probe position
        const int sbit = lo & 31;
        const long long w0 = (long long)lo >> 5;
        if (w0 + kGrainWindowWords > wbase + kSegStageWords) break;  // window not staged: refill
        // lane l < 24 owns bits [64 l, 64 l + 64) from bit 0 of word w0 (grain_seg.cuh)
        const int wi = (int)(w0 - wbase) + 2 * lane;
        const unsigned long long bits =
            lane < 24 ? ((unsigned long long)stage[wi] | ((unsigned long long)stage[wi + 1] << 32)) : 0ull;
        const SegSplit sp = seg_lane_split(bits, lane, sbit);
        const unsigned b_ge = __ballot_sync(0xffffffffu, sp.ge != 0ull);
        const unsigned b_lt = __ballot_sync(0xffffffffu, sp.lt != 0ull);
        const int lane_ge = b_ge ? __ffs((int)b_ge) - 1 : -1;
        const int lane_lt = b_lt ? 31 - __clz((int)b_lt) : -1;
        const unsigned long long w_ge = __shfl_sync(0xffffffffu, sp.ge, lane_ge & 31);
        const unsigned long long w_lt = __shfl_sync(0xffffffffu, sp.lt, lane_lt & 31);
        const int rel = seg_pick(lane_ge, w_ge, lane_lt, w_lt, sbit);
        int idx;
        if (rel >= 0) {
          idx = (int)(w0 * 32) + rel;
        } else {
          // no look-7 crossing within +-749: first look-3 crossing at or after start + 2250
          const long long s = (long long)start + kGrainPreferred + kGrainPreferred / 2;
          idx = -1;
          for (long long kb = s >> 5; kb < tr.nwords; kb += 32) {
            const long long kw = kb + lane;
            uint32_t word = kw < tr.nwords ? __ldg(tr.z3 + kw) : 0u;
            if (kw == (s >> 5)) word &= 0xffffffffu << (int)(s & 31);
            const unsigned any = __ballot_sync(0xffffffffu, word != 0u);
            if (any) {
              const int src = __ffs((int)any) - 1;
              const uint32_t wv = __shfl_sync(0xffffffffu, word, src);
              idx = (int)((kb + src) * 32 + (__ffs((int)wv) - 1));
              break;
            }
          }
          if (idx < 0) {
            done = 1;
            break;
          }
        }
        if (lane == 0 && count < cap) {
          tr.g_start[count] = start;
          tr.g_len[count] = idx - start;
        }
        ++count;
        start = idx;
      }
      __syncwarp();  // every lane has read the previous state before lane 0 replaces it
      if (lane == 0) {
        s_start = start;
        s_count = count;
        s_done = done;
      }
    }
    __syncthreads();
  }
  if (tid == 0) *tr.count = s_count;
}

cudaError_t launch_grain_segment(const GrainSegTrack* tracks_dev, int ntracks, long long max_words, int cap,
                                 cudaStream_t st) {
  if (ntracks <= 0) return cudaSuccess;
  if (max_words > 0) {
    dim3 grid((unsigned)((max_words + kSegWordsPerCta - 1) / kSegWordsPerCta), ntracks);
    grain_cross_kernel<<<grid, 256, 0, st>>>(tracks_dev);
  }
  grain_chain_kernel<<<ntracks, 256, 0, st>>>(tracks_dev, cap);
  return cudaGetLastError();
}


}  // namespace mlx
