/* include/melonix_gpu.h -- C ABI of libmelonix_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for melonix's one data-parallel hot path.  The reference
 * (mika314/melonix) has no FFI; its boundary is the C++ class surface `Spec` / `SpecCache`
 * (reference spec.hpp:11-39, spec-cache.hpp:13-39) plus the private resynthesis loop
 * App::process / App::exportWav (reference app.cpp:294-345, 1194-1215).  The host-side C++ mirror
 * of those classes (melonix_b200/host/) and every other binding (Python ctypes in melonix_b200/,
 * see INTEGRATION.md) reach the GPU only through the functions declared here.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; every function returns MLX_OK (0) or a
 *     negative mlx_status and leaves a message retrievable with mlx_last_error() (thread-local).
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x mlx_create fails.
 *   - calls on one mlx_ctx must be serialised by the caller (the host `Spec` calls from its single
 *     worker thread).  Work is issued on the context's stream (mlx_set_stream) and is asynchronous
 *     for the *_dev entry points until mlx_sync(); host-pointer entry points return when done.
 *   - "frame f" follows the reference's Spec job convention (spec.cpp:47, spec-cache.cpp:63-65):
 *     it covers samples [(f+1)*hop - fftN, (f+1)*hop), zero outside [0, n); F = ceil(n / hop).
 */
#ifndef MELONIX_GPU_H
#define MELONIX_GPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MLX_API __attribute__((visibility("default")))
#else
#define MLX_API
#endif

typedef enum {
  MLX_OK = 0,
  MLX_ERR_INVALID = -1,     /* bad argument                                      */
  MLX_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail)        */
  MLX_ERR_NOMEM = -3,       /* device or host allocation failed                   */
  MLX_ERR_STATE = -4,       /* call order violated (e.g. no tracks uploaded)       */
  MLX_ERR_UNSUPPORTED = -5  /* size / ratio outside what the kernels implement     */
} mlx_status;

typedef struct mlx_ctx mlx_ctx; /* opaque: device buffers, tables, stream, scratch */

/* ---- context ------------------------------------------------------------------------------ */
MLX_API int mlx_create(mlx_ctx **out, int device);
MLX_API void mlx_destroy(mlx_ctx *ctx);
MLX_API const char *mlx_last_error(void);
/* cuda_stream: a cudaStream_t (NULL = the legacy default stream). */
MLX_API int mlx_set_stream(mlx_ctx *ctx, void *cuda_stream);
MLX_API int mlx_sync(mlx_ctx *ctx);
/* sm_count, cc = major*10+minor, total device memory in bytes. Any pointer may be NULL. */
MLX_API int mlx_device_info(mlx_ctx *ctx, int *sm_count, int *cc, size_t *total_mem);
/* number of kernels launched by this context since creation (bench.py's gpu_launches). */
MLX_API int64_t mlx_launch_count(const mlx_ctx *ctx);

/* Per-kernel device timing with CUDA events on the context stream (what bench.py's roofline
 * object is computed from).  kinds: 0 = pv_analyze, 1 = pv_scan, 2 = pv_synth, 3 = spec, 4 = grain.
 * mlx_profile_read synchronises the stream, adds up the time between consecutive launch marks per
 * kind into ms[5] / launches[5], and clears the marks when reset != 0. */
MLX_API int mlx_profile_enable(mlx_ctx *ctx, int on);
MLX_API int mlx_profile_read(mlx_ctx *ctx, double *ms, int64_t *launches, int reset);

/* ---- tracks ------------------------------------------------------------------------------- */
/* Replaces Spec::Spec(std::span<float> wav) (reference spec.cpp:10-16): the reference keeps a
 * non-owning view of the mono track; here the samples are copied once into HBM, zero-padded on
 * both sides so that the reference's "outside [0,n) reads as 0" rule (spec.cpp:50-54) needs no
 * branches in the phase-vocoder kernels.  wav[t] points at n[t] floats (host memory). */
MLX_API int mlx_upload_tracks(mlx_ctx *ctx, const float *const *wav, const int64_t *n, int ntracks);
/* Same, from device memory (device-to-device copy on the context stream). */
MLX_API int mlx_upload_tracks_dev(mlx_ctx *ctx, const float *const *wav_dev, const int64_t *n, int ntracks);
MLX_API int mlx_num_tracks(const mlx_ctx *ctx);
MLX_API int64_t mlx_track_len(const mlx_ctx *ctx, int track);

/* ---- Spec path: STFT magnitude (replaces Spec::internalGetSpec, reference spec.cpp:44-66) ---- */
/* `count` jobs (start,end) exactly as passed to Spec::getSpec (spec.cpp:18): window [end-fftN,end),
 * samples before `start` decayed by expf(-2.5e-4f*(start-i)), outside [0,n) zero; output row j =
 * fftN/2 floats |FFT|/fftN (Nyquist bin dropped).  fftN in {512,...,32768}; the reference's
 * SpectrSize is 32768 (spec.cpp:8).  start_end = [count][2] int32, out = [count][fftN/2]. */
MLX_API int mlx_spec_batch(mlx_ctx *ctx, int track, int fftN, const int32_t *start_end, int count,
                   float *out);
MLX_API int mlx_spec_batch_dev(mlx_ctx *ctx, int track, int fftN, const int32_t *start_end_dev, int count,
                       float *out_dev);
/* Regular-hop jobs generated on the device: job f = (f*hop, (f+1)*hop), f in
 * [first_frame, first_frame+count).  out_dev = [count][fftN/2]. */
MLX_API int mlx_spec_frames_dev(mlx_ctx *ctx, int track, int fftN, int hop, int64_t first_frame,
                        int64_t count, float *out_dev);
/* The same for EVERY uploaded track in ONE launch (all F[t] = ceil(n[t] / hop) frames of track t; out_dev =
 * host array of ntracks device pointers, out_dev[t] = [F[t]][fftN/2]): enough CTAs to fill the chip where a
 * single track's launch lasts tens of microseconds. */
MLX_API int mlx_spec_frames_all_dev(mlx_ctx *ctx, int fftN, int hop, float *const *out_dev);
/* Spec + the colour ramp of SpecCache::populateTex fused (reference spec-cache.cpp:77-96):
 * out_rgb = [count][fftN/2][3] bytes, k = the brightness gain (spec-cache.cpp:79). */
MLX_API int mlx_spec_batch_rgb(mlx_ctx *ctx, int track, int fftN, const int32_t *start_end, int count,
                       float k, uint8_t *out_rgb);

/* ---- phase-vocoder path (NOT IN REFERENCE; PV-spec v1) --------------------------- */
typedef struct {
  int fftN;            /* 512, 1024, 2048, 4096 or 8192                                        */
  int hop;             /* must be fftN / 4                                                      */
  float rate;          /* pitch ratio, host powf(2.f, semitones/12.f) as reference app.cpp:297  */
  double sample_rate;  /* for f0 and the 50..2000 Hz peak-search band                           */
  /* optional per-track device arrays [F] of per-frame ratios (NULL, or entries NULL -> rate)   */
  const float *const *rate_per_frame_dev;
  /* time-range sharding (multi-GPU): only frames [frame_begin, frame_end) are owned by this
   * call; -1,-1 = all.  With frame_begin > 0 the track buffer must hold the halo frame
   * frame_begin-1 and output hops [frame_begin, frame_end) are written.                        */
  int64_t frame_begin, frame_end;
  /* accumulated synthesis phase carried in from earlier frames of the same track: per-track
   * device arrays of fftN/2+1 uint32 (NULL, or entries NULL -> 0).                              */
  const uint32_t *const *phase_in_dev;
  /* budget in MiB for the analysis->synthesis intermediates (8*(fftN/2+32) bytes per frame):
   * >0 tiles the frame range into waves of that size (e.g. to keep them in L2); <0 = one wave over
   * the whole range; 0 = library default (one wave unless that exceeds 1/4 of device memory).     */
  int wave_mib;
} mlx_pv_params;

/* Full pipeline on the uploaded tracks.  Per-track outputs (entries or whole arrays may be NULL):
 * out_wav[t]: n[t] floats; out_peak[t]: F[t] int32 peak bins; out_f0[t]: F[t] floats (Hz).
 * Only elements belonging to frames/hops in [frame_begin, frame_end) are written. */
MLX_API int mlx_pv_run(mlx_ctx *ctx, const mlx_pv_params *p, float *const *out_wav, int32_t *const *out_peak,
               float *const *out_f0);
MLX_API int mlx_pv_run_dev(mlx_ctx *ctx, const mlx_pv_params *p, float *const *out_wav_dev,
                   int32_t *const *out_peak_dev, float *const *out_f0_dev);
/* Analysis only over [frame_begin, frame_end): writes the per-track phase totals of the owned
 * frames (fftN/2+1 uint32 each, device) -- the quantity ranks exchange when one file is sharded
 * by time range -- plus peak / f0. */
MLX_API int mlx_pv_phase_totals_dev(mlx_ctx *ctx, const mlx_pv_params *p, uint32_t *const *totals_dev,
                            int32_t *const *out_peak_dev, float *const *out_f0_dev);
/* Split form of mlx_pv_run_dev, for callers that need the phase totals BEFORE synthesis (time-range
 * sharding: analyse once, exchange the totals, synthesise with the carried-in phase).
 * mlx_pv_analyze_dev runs the analysis (K_A) over [frame_begin, frame_end) exactly once, leaves the
 * shifted magnitudes and chunk-local phase sums staged in device memory (one wave: 8*(fftN/2+32) bytes
 * per frame and track), and writes the per-track phase totals of the owned frames (fftN/2+1 uint32,
 * device) plus peak / f0.  mlx_pv_synth_dev then applies p->phase_in_dev to the staged analysis (a
 * microsecond rescan of the chunk totals) and synthesises the owned hops.  Same p (geometry, rate,
 * frame range) in both calls; any upload or other PV call in between invalidates the staged analysis
 * (MLX_ERR_STATE).  analyze + synth is bit-identical to mlx_pv_run_dev with the same phase_in_dev. */
MLX_API int mlx_pv_analyze_dev(mlx_ctx *ctx, const mlx_pv_params *p, uint32_t *const *totals_dev,
                               int32_t *const *out_peak_dev, float *const *out_f0_dev);
MLX_API int mlx_pv_synth_dev(mlx_ctx *ctx, const mlx_pv_params *p, float *const *out_wav_dev);
/* Inspection of a staged analysis (after mlx_pv_analyze_dev): for frames [frame_begin, frame_begin+count)
 * of `track`, what the synthesis will be built from -- smag_dev[count][fftN/2+1] shifted magnitudes and
 * phase_dev[count][fftN/2+1] accumulated synthesis phases in 2^-32 turns (PV-spec A.5 / A.6: the oracle's
 * `smag` and `acc`), relative to the first analysed frame (a carried-in phase is added only at synthesis).
 * Used by the parity tests to compare every bin of every frame with the oracle, not only the audio. */
MLX_API int mlx_pv_stage_export_dev(mlx_ctx *ctx, int track, int64_t frame_begin, int64_t count, float *smag_dev,
                                    uint32_t *phase_dev);
/* End-to-end convenience for host buffers: uploads `wav`, runs, downloads, with the copies of
 * track t+1 / t-1 overlapped with the kernels of track t on separate streams. */
MLX_API int mlx_pv_process_host(mlx_ctx *ctx, const mlx_pv_params *p, const float *const *wav,
                        const int64_t *n, int ntracks, float *const *out_wav,
                        int32_t *const *out_peak, float *const *out_f0);

/* The same with a choice of sample format on each side of the PCIe wire.  MLX_FMT_I16 input: int16 PCM,
 * x = s / 32768 (exact; what the reference's decoder makes of 16-bit audio, swr s16 -> flt, app.cpp:640-690);
 * MLX_FMT_I16 output: int16(x * 32767.) by truncation, no clamp -- the conversion App::exportWav applies
 * before save-wav (reference app.cpp:1209-1212), fused into the synthesis kernel, so the bytes handed to
 * saveWav are produced on the device.  int16 on both sides halves the bytes per step.  wav[t] / out_wav[t]
 * point at n[t] samples of the stated format. */
typedef enum { MLX_FMT_F32 = 0, MLX_FMT_I16 = 1 } mlx_sample_format;
MLX_API int mlx_pv_process_host_fmt(mlx_ctx *ctx, const mlx_pv_params *p, const void *const *wav, int in_format,
                                    const int64_t *n, int ntracks, void *const *out_wav, int out_format,
                                    int32_t *const *out_peak, float *const *out_f0);

/* ---- one long file across the GPUs of a node: time-range sharding (BASELINE configs[3]) ---------
 * One rank (process or host thread) per GPU, each with its own mlx_ctx.  Rank r owns frames
 * [F*r/G, F*(r+1)/G) -- frame convention of reference spec.cpp:47 / spec-cache.cpp:63-65 -- and the
 * output hops of those frames.  The communicator wraps NCCL (bound at run time with
 * dlopen("libnccl.so.2"); MLX_NCCL_LIB overrides the name): rank 0 obtains 128 opaque bytes from
 * mlx_comm_unique_id, the caller distributes them (any transport), every rank calls mlx_comm_create. */
typedef struct mlx_comm mlx_comm;
typedef struct {
  int64_t frame_begin, frame_end; /* owned frames (global indices)                                   */
  int64_t own_lo, own_hi;         /* owned samples = output hops of the owned frames                  */
  int64_t need_lo, need_hi;       /* samples the rank must hold: owned + both seam overlaps           */
  int64_t frame_offset;           /* local frame index = global frame index - frame_offset            */
} mlx_time_shard;
/* Pure host function: the shard of `rank`.  Left overlap = fftN samples + the halo frame's hop (the
 * frame before the first owned one seeds the phase difference), right overlap = 3 hops (the tails of
 * the next three frames overlap-add into the owned hops); both clipped to [0, n). */
MLX_API int mlx_shard_frames(int64_t n, int fftN, int hop, int world, int rank, mlx_time_shard *out);
MLX_API int mlx_comm_unique_id(void *id128);
MLX_API int mlx_comm_create(mlx_comm **out, mlx_ctx *ctx, const void *id128, int world, int rank);
MLX_API void mlx_comm_destroy(mlx_comm *comm);
MLX_API int mlx_comm_info(const mlx_comm *comm, int *world, int *rank, int *nccl_version);
/* Phase-vocodes `ntracks` mono tracks of n_total samples each (the planar channels of one file) across
 * the ranks of `comm`.  own_dev[t]: this rank's owned samples [own_lo, own_hi) of track t (device).
 * Per call and rank: ONE NCCL group of send/recv of the overlap-region samples with the two
 * neighbours (received straight into the track buffer while the owned samples are copied in), ONE
 * analysis pass, ONE all-gather of the uint32 phase totals, synthesis of the owned hops.  Outputs
 * (device, entries or arrays may be NULL): out_own_dev[t] = own_hi-own_lo floats, peak_own_dev[t] /
 * f0_own_dev[t] = frame_end-frame_begin values.  Bit-identical to the unsharded mlx_pv_run_dev.
 * Collective: every rank of the communicator must call it with the same p, ntracks and n_total. */
MLX_API int mlx_pv_run_sharded_dev(mlx_ctx *ctx, mlx_comm *comm, const mlx_pv_params *p,
                                   const float *const *own_dev, int ntracks, int64_t n_total,
                                   float *const *out_own_dev, int32_t *const *peak_own_dev,
                                   float *const *f0_own_dev);

/* The same with host pointers (own[t], out_own[t]: own_hi-own_lo floats; peak_own[t] / f0_own[t]:
 * frame_end-frame_begin values): upload, run, download; returns when the results are in place.  This is
 * the call a C++ host makes, one thread (or process) per GPU. */
MLX_API int mlx_pv_run_sharded(mlx_ctx *ctx, mlx_comm *comm, const mlx_pv_params *p, const float *const *own,
                               int ntracks, int64_t n_total, float *const *out_own, int32_t *const *peak_own,
                               float *const *f0_own);

/* ---- grain path (replaces the inner loop of App::process, reference app.cpp:331-343, and the
 *      float->int16 conversion of App::exportWav, app.cpp:1209-1212) ---------------------------- */
/* The schedule (one row per App::process call that produced audio) is computed on the host by the
 * caller (melonix_b200/host/grain_schedule.hpp mirrors exportWav's cursor recurrence):
 *   g_start/g_len: the grain [start, start+len) in the track; g_rate: powf(2, pitchBend/12);
 *   out_off[ngrains+1]: output offset of each row (last = total rendered length);
 *   g_next: first sample of the grain that follows in output time (app.cpp:312-329).
 * out / out_i16 (either may be NULL): total_len = out_off[ngrains] + tail_zeros samples. */
MLX_API int mlx_grain_render(mlx_ctx *ctx, int track, const int32_t *g_start, const int32_t *g_len,
                     const float *g_rate, const int64_t *out_off, const float *g_next, int ngrains,
                     int tail_zeros, float *out, int16_t *out_i16);

/* ---- grain segmentation (replaces the zero-crossing search of App::preproc, reference
 *      app.cpp:156-235: probes start+1500 +0,+0,+1,-1,...,-749 with look-around 7, app.cpp:164-181,
 *      else forward scan from start+2250 with look-around 3, app.cpp:198-217) --------------------- */
/* Segments EVERY uploaded track in one call: the crossing predicates are evaluated for all samples
 * in parallel, the (serial, O(#grains)) chain runs one CTA per track.  Row t of g_start / g_len
 * receives the first min(counts[t], cap) grains of track t as (start, length) -- the key and span
 * size of the reference's `grains` map (app.hpp:40); counts[t] is the full grain count (call again
 * with a larger cap if it exceeds cap; n/751 + 1 always suffices).  Same results as
 * mlxh_grain_segment (include/melonix_host.h), bit for bit. */
MLX_API int mlx_grain_segment(mlx_ctx *ctx, int32_t *const *g_start, int32_t *const *g_len, int cap,
                      int32_t *counts);
/* Same with device pointers (g_start_dev / g_len_dev: host arrays of ntracks device row pointers,
 * counts_dev: device array of ntracks int32); asynchronous on the context stream. */
MLX_API int mlx_grain_segment_dev(mlx_ctx *ctx, int32_t *const *g_start_dev, int32_t *const *g_len_dev,
                          int cap, int32_t *counts_dev);

/* ---- waveform min/max pyramid (replaces App::calcPicks, reference app.cpp:347-378, and
 *      App::getMinMaxFromRange, app.cpp:380-426 -- the waveform display's level-of-detail cache) ---- */
/* Level l holds floor(n / 2^(l+1)) entries (min, max) over samples [i 2^(l+1), (i+1) 2^(l+1)); levels
 * exist while n > 2^(l+1), exactly the reference's `picks` (app.hpp:42).  All levels are stored back to
 * back as float pairs; mlx_picks_layout fills level_off[levels + 1] (first entry of each level, in
 * pairs) and returns the total number of pairs.  Both are pure host functions of n. */
MLX_API int mlx_picks_levels(int64_t n);
MLX_API int64_t mlx_picks_layout(int64_t n, int64_t *level_off);
/* Builds the pyramid of an uploaded track: pairs = [total][2] floats (host / device memory). */
MLX_API int mlx_picks_build(mlx_ctx *ctx, int track, float *pairs);
MLX_API int mlx_picks_build_dev(mlx_ctx *ctx, int track, float *pairs_dev);
/* Every uploaded track in one launch: pairs_dev = host array of ntracks device pointers. */
MLX_API int mlx_picks_build_all_dev(mlx_ctx *ctx, float *const *pairs_dev);
/* getMinMaxFromRange for `count` ranges (start, end) at once: out = [count][2] (min, max), including
 * the reference's behaviour at the edges (empty and out-of-range ranges give (0, 0) or the single
 * sample, app.cpp:382-396; the block that contains `start` is taken whole, app.cpp:399-408).  The
 * pyramid of `track` is built on first use and cached until the next upload. */
MLX_API int mlx_minmax_ranges(mlx_ctx *ctx, int track, const int32_t *start_end, int count, float *out);

#ifdef __cplusplus
}
#endif
#endif /* MELONIX_GPU_H */
