/* include/melonix_host.h -- C entry points of libmelonix_host.so: the host-side (serial, tiny)
 * parts of the grain path that stay on the CPU by design and feed mlx_grain_render().
 *
 * They mirror, argument for argument, what the untouched front-end computes around the hot loop:
 *   grain segmentation   App::preproc     reference app.cpp:156-235
 *   warp maps            time2Sample / time2PitchBend / sample2Time / duration  app.cpp:1020-1122
 *   export recurrence    App::exportWav + the bookkeeping half of App::process
 *                        reference app.cpp:1194-1207, 294-329
 *   waveform queries     App::getMinMaxFromRange  reference app.cpp:380-426 (on a GPU-built pyramid)
 * The per-sample resampling loop itself (app.cpp:331-343) and the float->int16 conversion
 * (app.cpp:1209-1212) run on the GPU (mlx_grain_render, include/melonix_gpu.h).
 */
#ifndef MELONIX_HOST_H
#define MELONIX_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MLXH_API __attribute__((visibility("default")))
#else
#define MLXH_API
#endif

typedef struct {
  int sample;
  double note, dTime, pitchBend;
} mlxh_marker; /* reference marker.hpp:4-19 */

/* Zero-crossing grain segmentation.  Writes up to cap (start, len) pairs; returns the grain count
 * (call again with a larger cap if it exceeds cap). */
MLXH_API int mlxh_grain_segment(const float *wav, int64_t n, int32_t *g_start, int32_t *g_len, int cap);

MLXH_API double mlxh_sample2time(const mlxh_marker *m, int nm, int sample_rate, int sample);
MLXH_API int mlxh_time2sample(const mlxh_marker *m, int nm, int sample_rate, double t);
MLXH_API double mlxh_duration(const mlxh_marker *m, int nm, int sample_rate, int64_t n);
MLXH_API float mlxh_time2pitchbend(const mlxh_marker *m, int nm, int sample_rate, int64_t n, double t);

/* Replays exportWav's cursor recurrence and emits the render schedule consumed by
 * mlx_grain_render: one row per process() call that produced audio.  out_off has rows+1 entries.
 * Returns the number of rows (or -(needed rows) if cap is too small).  *tail_zeros receives the
 * count of zeros the reference appends when it runs out of grains (1500, app.cpp:303-309). */
MLXH_API int mlxh_export_schedule(const float *wav, int64_t n, int sample_rate, const mlxh_marker *m, int nm,
                                  const int32_t *g_start, const int32_t *g_len, int ngrains, int32_t *s_gstart,
                                  int32_t *s_glen, float *s_rate, int64_t *s_out_off, float *s_next, int cap,
                                  int *tail_zeros);

/* Waveform level-of-detail cache, host half (melonix_b200/host/picks.hpp): the pyramid is built on the
 * GPU (mlx_picks_build); single range queries -- one per screen column per UI frame in the reference,
 * App::getMinMaxFromRange app.cpp:380-426 -- are answered on the host from the downloaded pyramid. */
MLXH_API int mlxh_picks_levels(int64_t n);
MLXH_API int64_t mlxh_picks_layout(int64_t n, int64_t *level_off /* [levels + 1] */);
MLXH_API void mlxh_minmax_ranges(const float *wav, int64_t n, const float *pairs /* [total][2] */,
                                 const int32_t *start_end, int count, float *out /* [count][2] */);

/* The colour ramp of SpecCache::populateTex (reference spec-cache.cpp:77-96) on the host: rgb =
 * [count][3].  New columns are coloured by the fused GPU epilogue (mlx_spec_batch_rgb); the drop-in
 * Spec uses this only to recolour a column whose floats are cached when the brightness changes. */
MLXH_API void mlxh_colour_ramp(const float *mag, int count, float k, uint8_t *rgb);

#ifdef __cplusplus
}
#endif
#endif
