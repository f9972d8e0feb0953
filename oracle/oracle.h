/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the melonix hot path (reference @ /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The shipped product path (melonix_b200/) never does and has no CPU fallback.
 *
 * Parity status:
 *   - mlxo_spec_*   : restates Spec::internalGetSpec (reference spec.cpp:44-66).  The reference has
 *                     no tests or golden vectors; pinned instead against the
 *                     reference's own spec.cpp compiled unmodified (oracle/_ref, see ref_spec.cpp)
 *                     and against the analytic KAT-1/KAT-2 values.
 *   - mlxo_colormap : restates SpecCache::populateTex colour ramp (reference spec-cache.cpp:77-96).
 *                     PINNED: the reference's spec-cache.cpp compiled unmodified (oracle/_ref/
 *                     libapp_ref.so, GL calls shimmed; the texels given to glTexImage1D are read back)
 *                     produces the same bytes in all three segments of the ramp.
 *   - mlxo_grain_*  : restates App::preproc grain segmentation (reference app.cpp:156-235),
 *                     App::process (app.cpp:294-345), App::exportWav (app.cpp:1194-1215) and the
 *                     marker warp maps (app.cpp:1020-1122).  PINNED: the reference's app.cpp and
 *                     save-wav.cpp compiled unmodified against no-op UI / audio / codec headers
 *                     (oracle/shim_app, driver oracle/ref_app.cpp) give identical grains, float
 *                     samples (bit patterns), int16 samples and warp-map values (tests/test_oracle.py).
 *   - mlxo_picks_* / mlxo_minmax_ranges : restate App::calcPicks / App::getMinMaxFromRange
 *                     (reference app.cpp:347-426).  PINNED the same way (bit patterns, NaN and signed
 *                     zeros included); brute-force min/max over aligned ranges as a second anchor.
 *   - mlxo_pv_*     : NOT IN REFERENCE.  Double-precision restatement of PV-spec v1 (DESIGN.md,
 *                     from SURVEY.md Appendix A).  parity unpinned by reference: self-consistency
 *                     target only.
 */
#ifndef MLXO_ORACLE_H
#define MLXO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- Spec STFT magnitude (reference spec.cpp:44-66, N parameterised; reference N = 32768) ---- */
/* One job: window [end-N, end), exp-decay before `start`, c2c double FFT, |.|/N, first N/2 bins. */
int mlxo_spec_frame(const float *wav, int64_t n, int start, int end, int N, float *out /*N/2*/);
/* count jobs, start_end = [count][2]; out = [count][N/2]; nthreads<=0 -> all cores (OpenMP). */
int mlxo_spec_batch(const float *wav, int64_t n, int N, const int32_t *start_end, int count,
                    float *out, int nthreads);

/* ---- colour ramp (reference spec-cache.cpp:77-96) ---- */
void mlxo_colormap(const float *spec, int count, float k, uint8_t *rgb /*count*3*/);

/* ---- phase vocoder, PV-spec v1 (NOT IN REFERENCE) ---- */
typedef struct {
  int N;            /* FFT size (power of two), hop = N/4                                     */
  int hop;          /* must be N/4                                                             */
  double fs;        /* sample rate for f0 / peak search band                                   */
  float rate;       /* constant pitch ratio (host powf(2.f, semis/12.f), cf. app.cpp:297)       */
  const float *rate_per_frame; /* optional [F]; NULL -> `rate`                                  */
} mlxo_pv_params;

/* Outputs may be NULL.  y[n] output audio; peak[F]; f0[F]; margin[F] = (m1-m2)/m1 of the top-two
 * magnitudes in the search band (for the bit-exact peak-bin test).
 * dbg_inc / dbg_smag: optional [F][N/2+1] dumps of the exact phase increments / shifted mags. */
int mlxo_pv_run(const float *x, int64_t n, const mlxo_pv_params *p, float *y, int32_t *peak,
                float *f0, double *margin, uint32_t *dbg_inc, float *dbg_smag);
/* ntracks equal-length tracks laid out [ntracks][n]; OpenMP over tracks. */
int mlxo_pv_run_batch(const float *x, int64_t n, int ntracks, const mlxo_pv_params *p, float *y,
                      int32_t *peak, float *f0, int nthreads);
int64_t mlxo_pv_num_frames(int64_t n, int hop);
/* exact gather range for output bin j: k in [klo,khi] with
 * trunc(float(k)*r)==j, or klo>khi when empty. */
void mlxo_pv_gather_range(int j, float r, int nbins, int *klo, int *khi);

/* ---- grains (reference app.cpp:156-235, 294-345, 1020-1122, 1194-1215) ---- */
typedef struct {
  int sample;
  double note, dTime, pitchBend;
} mlxo_marker; /* reference marker.hpp:4-19 */

/* Segmentation: returns number of grains; writes starts/lens up to cap. */
int mlxo_grain_segment(const float *wav, int64_t n, int32_t *g_start, int32_t *g_len, int cap);
double mlxo_sample2time(const mlxo_marker *m, int nm, int sampleRate, int val);
int mlxo_time2sample(const mlxo_marker *m, int nm, int sampleRate, double val);
double mlxo_duration(const mlxo_marker *m, int nm, int sampleRate, int64_t n);
float mlxo_time2pitchbend(const mlxo_marker *m, int nm, int sampleRate, int64_t n, double val);
/* exportWav restatement: drives process() until it returns 0.  Writes up to cap samples into
 * pcm (float) and pcm16; returns the output length (may exceed cap: call again with more room).
 * Optional schedule outputs (one row per process() call that produced audio):
 *   s_gstart, s_glen, s_rate, s_out_off, s_next (first sample of the grain that follows in output
 *   time, app.cpp:312-329), with *nsched rows (cap_sched capacity). */
int64_t mlxo_grain_export(const float *wav, int64_t n, int sampleRate, const mlxo_marker *m, int nm,
                          const int32_t *g_start, const int32_t *g_len, int ngrains, float *pcm,
                          int16_t *pcm16, int64_t cap, int32_t *s_gstart, int32_t *s_glen,
                          float *s_rate, int64_t *s_out_off, float *s_next, int *nsched,
                          int cap_sched);

/* ---- waveform min/max pyramid (reference app.cpp:347-378 calcPicks, :380-426 getMinMaxFromRange) ---- */
int mlxo_picks_levels(int64_t n);
/* level_off[levels + 1] in pairs; returns the total number of pairs */
int64_t mlxo_picks_layout(int64_t n, int64_t *level_off);
void mlxo_picks_build(const float *wav, int64_t n, float *pairs /*[total][2]*/, const int64_t *level_off);
void mlxo_minmax_ranges(const float *wav, int64_t n, const float *pairs, const int64_t *level_off,
                        const int32_t *start_end, int count, float *out /*[count][2]*/);

int mlxo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
