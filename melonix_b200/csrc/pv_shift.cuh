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

// ---- the same increment with a 32-bit per-bin constant.  base = r_fix * kh * 2^30 + 2^25 is A * 2^26 + 2^25
// (mod 2^64) with A = 16 * kh * r_fix, and the flip term -+r_fix * 2^32 is -+64 r_fix after the shift, so
//     inc = (A + ((d * r_fix + 2^25) >> 26) + flip term) mod 2^32,   A = 16 * kh * r_fix mod 2^32
// (arithmetic shift of the signed 64-bit product = floor, which is what bits 26..57 of the sum hold).
// An empty K_j reads the all-zero record with A = (j & 3) << 30.
MLX_HD uint32_t shift_inc_a(uint32_t A, int d, uint32_t mb, int r_fix) {
#ifdef __CUDA_ARCH__
  long long B;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(B) : "r"(d), "r"(r_fix), "l"(1LL << 25));
  const uint32_t q = __funnelshift_r((uint32_t)B, (uint32_t)((unsigned long long)B >> 32), 26);
#else
  const long long B = (long long)d * (long long)r_fix + (1LL << 25);
  const uint32_t q = (uint32_t)((unsigned long long)B >> 26);
#endif
  const uint32_t r64 = (uint32_t)r_fix << 6;  // (mod 2^32: r_fix reaches 2^28)
  const uint32_t adj = (uint32_t)((int)mb >> 31) & (d < 0 ? r64 : 0u - r64);
  return A + q + adj;
}

// ---- where the (mag, d) record of input bin k lives inside a frame buffer of the analysis kernels: the
// pair (k, NC - k), k <= NC/2, shares ONE 16-byte slot -- record of k in its first 8 bytes, of NC - k in the
// second -- so that the thread that analysed the pair writes both with one conflict-free 16-byte store.
// Returned: byte offset from the frame buffer's base; `padded`: slot index goes through fft_pad().
// (Alternating the halves with bit 3 of the slot index, plus one all-zero record per pair of banks, removes
// part of the two-way conflicts of the gather phase's 8-byte reads -- measured: 60 M fewer wavefronts of
// 2 500 M per step and no change in the kernel time, so the plain layout stays.)
MLX_HD uint32_t rec_offset(int k, int NC, bool padded) {
  const int lo = k <= NC / 2 ? k : NC - k;
  return 16u * (uint32_t)(padded ? fft_pad(lo) : lo) + (k <= NC / 2 ? 0u : 8u);
}

struct ShiftConstA {
  uint32_t off;  // byte offset of the source record (or of the all-zero record)
  uint32_t A;
};
MLX_HD ShiftConstA make_shift_const_a(int j, uint32_t kk, uint32_t r_fix, int NC, bool padded, uint32_t zero_off) {
  const int klo = (int)(kk & 0xffffu), khi = (int)(kk >> 16);
  ShiftConstA c;
  if (klo <= khi) {
    c.off = rec_offset(khi, NC, padded);
    c.A = ((uint32_t)khi * r_fix) << 4;
  } else {
    c.off = zero_off;
    c.A = ((uint32_t)j & 3u) << 30;
  }
  return c;
}

}  // namespace mlx
