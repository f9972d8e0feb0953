/* oracle/pv_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * NOT IN REFERENCE: melonix contains no phase vocoder.  This is the
 * double-precision restatement of PV-spec v1 (SURVEY.md Appendix A with step 3
 * evaluated as d = arg(X_f * conj(X_{f-1}) * (-i)^k), which is the same wrapped phase difference
 * without an explicit rint()).  parity unpinned by reference -- self-consistency target only.
 *
 * Frame convention follows the reference's Spec jobs (spec.cpp:47, spec-cache.cpp:63-65):
 * frame f covers samples [(f+1)H - N, (f+1)H), zero outside [0,n); F = ceil(n/H).
 */
#include "fft64.h"
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int64_t mlxo_pv_num_frames(int64_t n, int hop) { return (n + hop - 1) / hop; }

/* A.5: K_j = { k : trunc(float(k) * r) == j }, evaluated with a float multiply exactly as the
 * GPU does.  Fills klo/khi for j in [0,nbins); empty ranges get klo=1, khi=0. */
static void gather_map(float r, int nbins, int *klo, int *khi) {
  for (int j = 0; j < nbins; ++j) {
    klo[j] = 1;
    khi[j] = 0;
  }
  for (int k = 0; k < nbins; ++k) {
    const float t = (float)k * r;
    const int j = (int)truncf(t);
    if (j < 0 || j >= nbins) continue;
    if (klo[j] > khi[j]) klo[j] = k;
    khi[j] = k;
  }
}

void mlxo_pv_gather_range(int j, float r, int nbins, int *klo, int *khi) {
  int *a = (int *)malloc(sizeof(int) * (size_t)nbins * 2);
  gather_map(r, nbins, a, a + nbins);
  *klo = a[j];
  *khi = a[nbins + j];
  free(a);
}

int mlxo_pv_run(const float *x, int64_t n, const mlxo_pv_params *p, float *y, int32_t *peak,
                float *f0, double *margin, uint32_t *dbg_inc, float *dbg_smag) {
  const int N = p->N, H = p->hop;
  if (N < 16 || (N & (N - 1)) || H * 4 != N) return -1;
  const int M = N / 2, nb = M + 1, osamp = N / H;
  const int64_t F = mlxo_pv_num_frames(n, H);
  mlxo_fft_plan *plan = mlxo_fft_plan_create(M);
  if (!plan) return -1;

  /* A.1 window: periodic Hann, computed in double, stored as float */
  float *w = (float *)malloc(sizeof(float) * (size_t)N);
  double sw2 = 0.0;
  for (int j = 0; j < N; ++j) {
    w[j] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)j / (double)N));
    sw2 += (double)w[j] * (double)w[j];
  }
  const float g = (float)((double)H / sw2); /* A.7, = 2/3 for Hann at 4x overlap */

  /* A.4 search band */
  int kmin = (int)ceil(50.0 * N / p->fs), kmax = (int)floor(2000.0 * N / p->fs);
  if (kmin < 1) kmin = 1;
  if (kmax > M) kmax = M;
  if (kmax < kmin) kmax = kmin;

  double *xw = (double *)malloc(sizeof(double) * (size_t)N);
  double *X = (double *)malloc(sizeof(double) * 2 * (size_t)nb);
  double *Xp = (double *)malloc(sizeof(double) * 2 * (size_t)nb);
  double *Y = (double *)malloc(sizeof(double) * 2 * (size_t)nb);
  double *yt = (double *)malloc(sizeof(double) * (size_t)N);
  double *sc = (double *)malloc(sizeof(double) * 2 * (size_t)N);
  double *mag = (double *)malloc(sizeof(double) * (size_t)nb);
  double *nu = (double *)malloc(sizeof(double) * (size_t)nb);
  uint32_t *acc = (uint32_t *)calloc((size_t)nb, sizeof(uint32_t));
  int *klo = (int *)malloc(sizeof(int) * (size_t)nb * 2), *khi = klo + nb;
  double *out = y ? (double *)calloc((size_t)n, sizeof(double)) : NULL;
  for (int k = 0; k < nb; ++k) { /* phi_{-1} = 0  <=>  X_{-1} = 1 */
    Xp[2 * k] = 1.0;
    Xp[2 * k + 1] = 0.0;
  }
  float rlast = NAN;

  for (int64_t f = 0; f < F; ++f) {
    const float r = p->rate_per_frame ? p->rate_per_frame[f] : p->rate;
    if (!(r == rlast)) {
      gather_map(r, nb, klo, khi);
      rlast = r;
    }
    const int64_t s0 = (f + 1) * (int64_t)H - N;
    for (int m = 0; m < N; ++m) {
      const int64_t i = s0 + m;
      xw[m] = (i < 0 || i >= n) ? 0.0 : (double)w[m] * (double)x[i]; /* A.2 */
    }
    mlxo_rfft(plan, xw, X, sc);
    for (int k = 0; k < nb; ++k) {
      const double a = X[2 * k], b = X[2 * k + 1], c = Xp[2 * k], d = Xp[2 * k + 1];
      mag[k] = sqrt(a * a + b * b);
      double zr = a * c + b * d, zi = b * c - a * d; /* X conj(Xp) */
      double t;
      switch (k & 3) { /* times (-i)^k = e^{-2 pi i k/osamp}, osamp = 4 */
        case 1: t = zr; zr = zi; zi = -t; break;
        case 2: zr = -zr; zi = -zi; break;
        case 3: t = zr; zr = -zi; zi = t; break;
        default: break;
      }
      /* A.3 with the silence gate of PV-spec v1: |Z| <= 1e-18 -> d = 0 */
      /* k = 0 and k = N/2 are purely real bins: Im Z is defined as +0 there, so d is 0 or +pi */
      if (k == 0 || k == M) zi = 0.0;
      const double dd = (zr * zr + zi * zi <= 1e-36) ? 0.0 : atan2(zi, zr);
      nu[k] = (double)k + (double)osamp * dd / (2.0 * M_PI);
    }
    memcpy(Xp, X, sizeof(double) * 2 * (size_t)nb);

    /* A.4 peak bin: lowest k on exact ties */
    int pk = kmin;
    for (int k = kmin + 1; k <= kmax; ++k)
      if (mag[k] > mag[pk]) pk = k;
    if (peak) peak[f] = pk;
    if (f0) f0[f] = (float)(nu[pk] * p->fs / (double)N);
    if (margin) {
      double m2 = 0.0;
      for (int k = kmin; k <= kmax; ++k)
        if (k != pk && mag[k] > m2) m2 = mag[k];
      margin[f] = mag[pk] > 0.0 ? (mag[pk] - m2) / mag[pk] : 0.0;
    }

    /* A.5 gather, A.6 exact phase accumulation, A.7 synthesis spectrum */
    for (int j = 0; j < nb; ++j) {
      double smag = 0.0, snu = (double)j;
      if (klo[j] <= khi[j]) {
        for (int k = klo[j]; k <= khi[j]; ++k) smag += mag[k];
        snu = (double)r * nu[khi[j]];
      }
      double t = snu / (double)osamp;
      t -= floor(t);
      const uint32_t inc = (uint32_t)(uint64_t)llrint(t * 4294967296.0);
      acc[j] += inc;
      if (dbg_inc) dbg_inc[(size_t)f * nb + j] = inc;
      if (dbg_smag) dbg_smag[(size_t)f * nb + j] = (float)smag;
      const double th = (double)acc[j] * (2.0 * M_PI / 4294967296.0);
      Y[2 * j] = smag * cos(th);
      Y[2 * j + 1] = smag * sin(th);
    }
    Y[1] = 0.0;
    Y[2 * M + 1] = 0.0;
    if (out) {
      mlxo_irfft(plan, Y, yt, sc);
      for (int m = 0; m < N; ++m) {
        const int64_t i = s0 + m;
        if (i < 0 || i >= n) continue;
        out[i] += (double)g * (double)w[m] * yt[m]; /* ascending-f summation order */
      }
    }
  }
  if (out) {
    for (int64_t i = 0; i < n; ++i) y[i] = (float)out[i];
    free(out);
  }
  free(w); free(xw); free(X); free(Xp); free(Y); free(yt); free(sc); free(mag); free(nu);
  free(acc); free(klo);
  mlxo_fft_plan_destroy(plan);
  return 0;
}

int mlxo_pv_run_batch(const float *x, int64_t n, int ntracks, const mlxo_pv_params *p, float *y,
                      int32_t *peak, float *f0, int nthreads) {
  const int64_t F = mlxo_pv_num_frames(n, p->hop);
  int rc = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int t = 0; t < ntracks; ++t) {
    const int r = mlxo_pv_run(x + (size_t)t * n, n, p, y ? y + (size_t)t * n : NULL,
                              peak ? peak + (size_t)t * F : NULL, f0 ? f0 + (size_t)t * F : NULL,
                              NULL, NULL, NULL);
    if (r) rc = r;
  }
  return rc;
}
