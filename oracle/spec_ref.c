/* oracle/spec_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Restates Spec::internalGetSpec, reference spec.cpp:44-66, with the FFT size as a parameter
 * (the reference hard-codes SpectrSize = 8*4096, spec.cpp:8). */
#include "fft64.h"
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int mlxo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static void spec_one(const mlxo_fft_plan *plan, const float *wav, int64_t n, int start, int end,
                     int N, float *out, double *in, double *outc, double *scratch) {
  /* spec.cpp:46-59: fill N complex doubles */
  int p = 0;
  for (int64_t i = (int64_t)end - N; i < end; ++i, ++p) {
    in[2 * p + 1] = 0.0;
    if (i >= n || i < 0) {
      in[2 * p] = 0.0;
      continue;
    }
    if (i >= start)
      in[2 * p] = wav[i];
    else /* float product, then widened (spec.cpp:58) */
      in[2 * p] = (double)(expf(-2.5e-4f * (float)(start - i)) * wav[i]);
  }
  mlxo_fft_c2c(plan, in, outc, scratch, -1); /* spec.cpp:60 */
  for (int i = 0; i < N / 2; ++i)            /* spec.cpp:61-65 */
    out[i] = (float)(sqrt(outc[2 * i] * outc[2 * i] + outc[2 * i + 1] * outc[2 * i + 1]) / N);
}

int mlxo_spec_frame(const float *wav, int64_t n, int start, int end, int N, float *out) {
  return mlxo_spec_batch(wav, n, N, (const int32_t[]){start, end}, 1, out, 1);
}

int mlxo_spec_batch(const float *wav, int64_t n, int N, const int32_t *start_end, int count,
                    float *out, int nthreads) {
  mlxo_fft_plan *plan = mlxo_fft_plan_create(N);
  if (!plan) return -1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    double *in = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *oc = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *sc = (double *)malloc(sizeof(double) * 2 * (size_t)N);
#pragma omp for schedule(static)
    for (int j = 0; j < count; ++j)
      spec_one(plan, wav, n, start_end[2 * j], start_end[2 * j + 1], N, out + (size_t)j * (N / 2),
               in, oc, sc);
    free(in);
    free(oc);
    free(sc);
  }
  mlxo_fft_plan_destroy(plan);
  return 0;
}
