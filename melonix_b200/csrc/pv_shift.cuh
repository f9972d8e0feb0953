// melonix_b200/csrc/pv_shift.cuh -- exact phase increment of the bin shift (PV-spec A.5 / A.6) with
// per-bin launch constants.  Host/device so that tests/host/pv_shift_emul.cpp checks the ring
// arithmetic against a 128-bit evaluation of the defining formula without a GPU.
#pragma once
#include <stdint.h>

#include "fft.cuh"  // MLX_HD, fft_pad

namespace mlx {

// what the pair phase leaves for the gather phase, stored in the first 8 bytes of the bin's (dead)
// FFT slot: |X| with the cut-flip flag in its sign bit, and the wrapped phase advance in 2^-32 turns
struct MagD {
  float mag;
  int d;
};

// Frame-invariant part of the bin shift of one output bin j at a constant rate (MLX_GATHER_V2).
// With K_j = [klo, khi], kh = khi and d' the signed phase advance (d32, -+2^32 on a cut flip)
//     inc = (r_fix * (kh * 2^30 + d') + 2^25) >> 26   (mod 2^32)
// is a product in the ring of integers mod 2^64, so the kh term is added once per launch:
//     base = r_fix * kh * 2^30 + 2^25,   inc = (base + r_fix * d') >> 26.
// An empty K_j (s_nu = j, inc = frac(j / 4) * 2^32, smag = 0) is the same formula with
// base = (j & 3) << 56 (+ 2^25, which the shift drops) reading an all-zero (mag, d) record.  The
// result is bit-identical to shift_one_bin(); the per-frame work is one 32 x 32 + 64 multiply-add,
// the flip term and the shift.
struct ShiftConst {
  uint32_t slot;  // padded index of bin kh's (mag, d) record inside a frame buffer, or the zero record
  unsigned long long base;
};
MLX_HD ShiftConst make_shift_const(int j, uint32_t kk, uint32_t r_fix, int zero_slot) {
  const int klo = (int)(kk & 0xffffu), khi = (int)(kk >> 16);
  ShiftConst c;
  if (klo <= khi) {
    c.slot = (uint32_t)fft_pad(khi);
    c.base = (unsigned long long)r_fix * ((unsigned long long)khi << 30) + (1ULL << 25);
  } else {  // reads (mag, d) = (0, 0): smag = 0 and the product term vanishes
    c.slot = (uint32_t)zero_slot;
    c.base = ((unsigned long long)((uint32_t)j & 3u) << 56) + (1ULL << 25);
  }
  return c;
}

// inc of one frame from the launch constant and the bin's record: mb = bit pattern of the stored
// magnitude (sign bit = cut flip: d' = d + 2^32 when d < 0, d - 2^32 otherwise)
MLX_HD uint32_t shift_inc(unsigned long long base, int d, uint32_t mb, int r_fix) {
  const int hi_adj = ((int)mb >> 31) & (d < 0 ? r_fix : -r_fix);
  // base + d * r_fix as ONE signed 32 x 32 + 64 multiply-add (the compiler expands the C expression
  // into a 64 x 64 product), the flip term goes into the high word, the shift is a funnel shift
  unsigned long long prod;
#ifdef __CUDA_ARCH__
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(prod) : "r"(d), "r"(r_fix), "l"(base));
  const uint32_t lo = (uint32_t)prod, hi = (uint32_t)(prod >> 32) + (uint32_t)hi_adj;
  return __funnelshift_r(lo, hi, 26);
#else
  prod = base + (unsigned long long)((long long)d * (long long)r_fix);
  const uint32_t lo = (uint32_t)prod, hi = (uint32_t)(prod >> 32) + (uint32_t)hi_adj;
  return (uint32_t)((((unsigned long long)hi << 32) | lo) >> 26);
#endif
}

}  // namespace mlx
