/* oracle/fft64.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). See fft64.h. */
#include "fft64.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct mlxo_fft_plan {
  int n;
  double *w;  /* w[2k],w[2k+1] = cos,-sin(2 pi k / n), k in [0,n)  (forward twiddles)   */
  double *wr; /* real-transform twiddles for length 2n: cos,-sin(2 pi k /(2n)), k in [0,n] */
};

static void fill_twiddle(double *w, int count, int period) {
  for (int k = 0; k < count; ++k) {
    /* exact octant symmetry is unnecessary: glibc sin/cos are < 1 ulp */
    const double a = 2.0 * M_PI * (double)k / (double)period;
    w[2 * k] = cos(a);
    w[2 * k + 1] = -sin(a);
  }
}

mlxo_fft_plan *mlxo_fft_plan_create(int n) {
  if (n < 1 || (n & (n - 1))) return NULL;
  mlxo_fft_plan *p = (mlxo_fft_plan *)calloc(1, sizeof(*p));
  p->n = n;
  p->w = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  p->wr = (double *)malloc(sizeof(double) * 2 * ((size_t)n + 1));
  fill_twiddle(p->w, n, n);
  fill_twiddle(p->wr, n + 1, 2 * n);
  return p;
}

void mlxo_fft_plan_destroy(mlxo_fft_plan *p) {
  if (!p) return;
  free(p->w);
  free(p->wr);
  free(p);
}

int mlxo_fft_plan_size(const mlxo_fft_plan *p) { return p->n; }

/* Stockham autosort, decimation in time, radix 4 with one radix-2 stage when log2(n) is odd.
 * Stage with Ns = product of earlier radices:  butterfly j reads src[j + r*n/R], multiplies by
 * exp(dir*2*pi*i*k*r/(Ns*R)) with k = j mod Ns, and writes dst[(j-k)*R + k + r*Ns]. */
void mlxo_fft_c2c(const mlxo_fft_plan *p, const double *in, double *out, double *scratch, int dir) {
  const int n = p->n;
  const double *w = p->w;
  double *a = scratch, *b = scratch + 2 * (size_t)n;
  /* scratch must hold 4n doubles when in/out ping-pong is needed; we use out as one buffer. */
  (void)b;
  if (n == 1) {
    out[0] = in[0];
    out[1] = in[1];
    return;
  }
  /* count stages to land the final result in `out` */
  int lg = 0;
  while ((1 << lg) < n) ++lg;
  const int nstages = lg / 2 + (lg & 1);
  double *src, *dst;
  /* ping-pong between `a` (scratch) and `out`; arrange so the last write goes to out */
  memcpy(a, in, sizeof(double) * 2 * (size_t)n);
  src = a;
  dst = out;
  if ((nstages & 1) == 0) {
    /* even number of stages: first write must go to ... a -> out -> a -> out ; with src=a the
       sequence of dst is out,a,out,a: ends in a for even. Start from out instead. */
    memcpy(out, in, sizeof(double) * 2 * (size_t)n);
    src = out;
    dst = a;
  }
  const double sgn = (dir < 0) ? 1.0 : -1.0; /* table holds forward twiddles; conj for backward */
  int Ns = 1;
  int first_radix2 = lg & 1;
  while (Ns < n) {
    if (first_radix2) {
      first_radix2 = 0;
      const int t = n / 2;
      for (int j = 0; j < t; ++j) {
        const int k = j & (Ns - 1);
        const int tw = k * (n / (Ns * 2));
        const double wr_ = w[2 * tw], wi_ = sgn * w[2 * tw + 1];
        const double ar = src[2 * j], ai = src[2 * j + 1];
        const double xr = src[2 * (j + t)], xi = src[2 * (j + t) + 1];
        const double br = xr * wr_ - xi * wi_, bi = xr * wi_ + xi * wr_;
        const int j0 = ((j - k) << 1) + k;
        dst[2 * j0] = ar + br;
        dst[2 * j0 + 1] = ai + bi;
        dst[2 * (j0 + Ns)] = ar - br;
        dst[2 * (j0 + Ns) + 1] = ai - bi;
      }
      Ns *= 2;
    } else {
      const int t = n / 4;
      const int step = n / (Ns * 4);
      for (int j = 0; j < t; ++j) {
        const int k = j & (Ns - 1);
        const int t1 = k * step, t2 = 2 * t1, t3 = 3 * t1;
        const double w1r = w[2 * t1], w1i = sgn * w[2 * t1 + 1];
        const double w2r = w[2 * t2], w2i = sgn * w[2 * t2 + 1];
        const double w3r = w[2 * t3], w3i = sgn * w[2 * t3 + 1];
        const double x0r = src[2 * j], x0i = src[2 * j + 1];
        double yr = src[2 * (j + t)], yi = src[2 * (j + t) + 1];
        const double x1r = yr * w1r - yi * w1i, x1i = yr * w1i + yi * w1r;
        yr = src[2 * (j + 2 * t)];
        yi = src[2 * (j + 2 * t) + 1];
        const double x2r = yr * w2r - yi * w2i, x2i = yr * w2i + yi * w2r;
        yr = src[2 * (j + 3 * t)];
        yi = src[2 * (j + 3 * t) + 1];
        const double x3r = yr * w3r - yi * w3i, x3i = yr * w3i + yi * w3r;
        const double s02r = x0r + x2r, s02i = x0i + x2i, d02r = x0r - x2r, d02i = x0i - x2i;
        const double s13r = x1r + x3r, s13i = x1i + x3i, d13r = x1r - x3r, d13i = x1i - x3i;
        /* forward: -i*(d13) = (d13i, -d13r); backward: +i*(d13) = (-d13i, d13r) */
        const double jr = sgn * d13i, ji = -sgn * d13r;
        const int j0 = ((j - k) << 2) + k;
        dst[2 * j0] = s02r + s13r;
        dst[2 * j0 + 1] = s02i + s13i;
        dst[2 * (j0 + Ns)] = d02r + jr;
        dst[2 * (j0 + Ns) + 1] = d02i + ji;
        dst[2 * (j0 + 2 * Ns)] = s02r - s13r;
        dst[2 * (j0 + 2 * Ns) + 1] = s02i - s13i;
        dst[2 * (j0 + 3 * Ns)] = d02r - jr;
        dst[2 * (j0 + 3 * Ns) + 1] = d02i - ji;
      }
      Ns *= 4;
    }
    double *tmp = src;
    src = dst;
    dst = tmp;
  }
  /* result is in src; by construction src == out */
  if (src != out) memcpy(out, src, sizeof(double) * 2 * (size_t)n);
}

void mlxo_rfft(const mlxo_fft_plan *p, const double *x, double *X, double *scratch) {
  const int M = p->n; /* N/2 */
  double *Z = scratch;             /* 2M doubles */
  double *tmp = scratch + 2 * (size_t)M; /* 2M doubles */
  /* z[m] = x[2m] + i x[2m+1]: x is already that layout */
  mlxo_fft_c2c(p, x, Z, tmp, -1);
  const double *wr = p->wr;
  for (int k = 0; k <= M; ++k) {
    const int k1 = (k == M) ? 0 : k;
    const int k2 = (k == 0) ? 0 : M - k;
    const double ar = Z[2 * k1], ai = Z[2 * k1 + 1];
    const double br = Z[2 * k2], bi = -Z[2 * k2 + 1]; /* conj(Z[M-k]) */
    const double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
    const double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
    /* X = E - i * W^k * D */
    const double c = wr[2 * k], s = wr[2 * k + 1]; /* W^k = c + i s  (s = -sin) */
    const double tr = dr * c - di * s, ti = dr * s + di * c;
    X[2 * k] = er + ti;
    X[2 * k + 1] = ei - tr;
  }
}

void mlxo_irfft(const mlxo_fft_plan *p, const double *X, double *y, double *scratch) {
  const int M = p->n;
  double *Z = scratch;
  double *tmp = scratch + 2 * (size_t)M;
  const double *wr = p->wr;
  for (int k = 0; k < M; ++k) {
    const double ar = X[2 * k], ai = X[2 * k + 1];
    const double br = X[2 * (M - k)], bi = -X[2 * (M - k) + 1]; /* conj(X[M-k]) */
    const double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
    const double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
    /* O = D * conj(W^k);  Z = E + i O */
    const double c = wr[2 * k], s = -wr[2 * k + 1];
    const double or_ = dr * c - di * s, oi = dr * s + di * c;
    Z[2 * k] = er - oi;
    Z[2 * k + 1] = ei + or_;
  }
  mlxo_fft_c2c(p, Z, y, tmp, +1);
  const double inv = 1.0 / (double)M;
  for (int i = 0; i < 2 * M; ++i) y[i] *= inv;
}
