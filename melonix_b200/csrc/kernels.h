// melonix_b200/csrc/kernels.h -- internal launch interface between capi.cu and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft.cuh"

namespace mlx {

// Zero padding kept in HBM around every uploaded track so that frames / TMA tiles that reach
// outside [0, n) read zeros without branches (reference rule: spec.cpp:50-54).
constexpr int kPadFront = 8192;   // >= max fftN of the phase-vocoder path
constexpr int kPadBack = 16384;   // >= fftN + one batch of hops
constexpr int kPvMaxN = 8192;

struct PvTrack {
  const float* x;          // sample 0 of the padded device copy
  long long n;             // samples
  long long F;             // frames = ceil(n / hop)
  float* out;              // [n] or nullptr
  short* out16;            // [n] or nullptr: the output as the reference's export sink takes it,
                           // int16(x * 32767.) by truncation, no clamp (app.cpp:1209-1212)
  int* peak;               // [F] or nullptr
  float* f0;               // [F] or nullptr
  const float* rate_pf;    // [F] or nullptr
};

// Twiddle / window tables for one fftN (device pointers, built on the host in double).
struct PvTables {
  const cplx<double>* tw_d;   // exp(-2 pi i m / NC), m in [0, NC)       NC = fftN/2
  const cplx<double>* twr_d;  // exp(-2 pi i k / fftN), k in [0, NC/2]
  const cplx<float>* tw_f;
  const cplx<float>* twr_f;
  const float* win;           // periodic Hann, float, [fftN]
  const double* win_d;        // the same float values widened to double (exact), [fftN]
  const float* wsyn;          // gain * win / fftN  (synthesis window incl. irfft 1/N), [fftN]
  const cplx<double>* tw1_d;  // exp(-2 pi i k r / 256) at [(r - 1) * 16 + k], r = 1..15, k < 16 (fft.cuh: twiddle_table16)
};

// One wave = owned frame window [wb, we) of every track; intermediates live in `rows` scratch rows
// per track (rows = we - wb + 3: the three frames after `we` are analysed again for overlap-add).
struct PvWave {
  long long wb, we;
  int rows;
  int CA, nchunksA;  // analysis chunk (frames per CTA) and chunks per track
  int CS, nchunksS;  // synthesis chunk (hops per CTA)
  float rate;
  float fs_over_N;
  int kmin, kmax;
  // bin-shift table for the constant `rate` (PV-spec A.5), built on the host per call:
  //   gk[j] = klo | khi << 16 with K_j = [klo, khi] (klo = 1, khi = 0 when empty)
  const uint32_t* gk;
  long long r_fix;  // rate * 2^26, exact (a float has 24 significant bits)
};

struct PvScratch {
  uint2* stage;     // [ntracks][rows][NBP]  per output bin: .x = shifted magnitude (float bits), .y = chunk-local
                    //                       inclusive phase sum -- one 8-byte record, one store in K_A, one load in K_S
  uint32_t* tot;    // [ntracks][nchunksA][NBP] chunk totals over all frames of the chunk
  uint32_t* totc;   // [ntracks][nchunksA][NBP] chunk totals over its frames < we (carry to the next wave)
  uint32_t* pre;    // [ntracks][nchunksA][NBP] exclusive prefix (incl. carry)
  uint32_t* carry;  // [ntracks][NBP] running phase at frame wb (updated to `we` by the scan)
};

constexpr int pv_nbp(int fftN) { return fftN / 2 + 32; }
int pv_group_count(int fftN);          // frames per batch of the synthesis kernel
int pv_group_count_analyze(int fftN);  // frames per batch of the analysis kernel
int pv_threads(int fftN);
size_t pv_analyze_smem(int fftN);
size_t pv_synth_smem(int fftN);

// constant_rate_up: no track has per-frame ratios and the ratio is >= 1 (selects the specialised instantiation)
cudaError_t launch_pv_analyze(int fftN, const PvTrack* tracks_dev, int ntracks, const PvWave& wv,
                              const PvTables& tb, const PvScratch& sc, bool constant_rate_up, cudaStream_t st);
cudaError_t launch_pv_scan(int fftN, int ntracks, const PvWave& wv, const PvScratch& sc,
                           cudaStream_t st);
// out16: write PvTrack::out16 (int16) instead of PvTrack::out (float)
cudaError_t launch_pv_synth(int fftN, const PvTrack* tracks_dev, int ntracks, const PvWave& wv,
                            const PvTables& tb, const PvScratch& sc, bool out16, cudaStream_t st);
// int16 PCM -> float samples x = s / 32768 (exact; what the reference's decoder hands to App, swr s16 -> flt)
cudaError_t launch_pcm16_to_float(const short* in, float* out, long long n, cudaStream_t st);
cudaError_t pv_configure(int fftN);  // cudaFuncSetAttribute for the instantiation
// staged analysis of one track -> per frame and bin: shifted magnitude and ABSOLUTE accumulated synthesis
// phase (chunk prefix + chunk-local sum), frames [f0, f0 + count) of the staged wave
cudaError_t launch_pv_stage_export(int fftN, int track, const PvWave& wv, const PvScratch& sc, long long f0,
                                   long long count, float* smag, uint32_t* phase, cudaStream_t st);
// K_A2 (pv_analyze2.cu): constant rate >= 1, fftN in {1024, 2048}; bit-identical to launch_pv_analyze
bool pv_analyze2_supported(int fftN);
int pv_analyze2_band_capacity(int fftN);  // largest kmax - kmin + 1 its peak search holds
int pv_analyze2_frames_per_batch(int fftN);
cudaError_t pv_analyze2_configure(int fftN);
cudaError_t launch_pv_analyze2(int fftN, const PvTrack* tracks_dev, int ntracks, const PvWave& wv,
                               const PvTables& tb, const PvScratch& sc, cudaStream_t st);

// ---- Spec
// per-track part of a batched regular-hop launch (blockIdx.y = track): overrides x / n / count / out / rgb
struct SpecTrackDesc {
  const float* x;
  long long n;
  long long count;
  float* out;
  unsigned char* rgb;
};
struct SpecArgs {
  const float* x;       // sample 0 of the padded device copy
  long long n;
  const int* jobs;      // [count][2] (start,end) or nullptr -> regular hop
  int hop;
  long long first_frame;
  long long count;
  float* out;           // [count][fftN/2] or nullptr
  unsigned char* rgb;   // [count][fftN/2][3] or nullptr
  float kcol;
  const cplx<float>* tw_f;
  const cplx<float>* twr_f;
  const float* decay;   // decay[d] = expf(-2.5e-4f * d), d in [0, fftN] (host glibc expf, spec.cpp:58)
  const SpecTrackDesc* multi;  // nullptr, or [ntracks] (device): one launch over every track, regular hop only
  int ntracks;
};
cudaError_t spec_configure(int fftN);
cudaError_t launch_spec(int fftN, const SpecArgs& a, cudaStream_t st);

// ---- grains
struct GrainArgs {
  const float* x;
  const int* g_start;
  const int* g_len;
  const float* g_rate;
  const long long* out_off;  // [ngrains + 1]
  const float* g_next;
  int ngrains;
  long long total;           // out_off[ngrains] + tail zeros
  float* out;
  short* out_i16;
};
cudaError_t launch_grain(const GrainArgs& a, cudaStream_t st);

// ---- grain segmentation (K8)
struct GrainSegTrack {
  const float* x;      // sample 0 of the padded device copy
  long long n;
  long long nwords;    // words of z7 / z3: ceil(n / 32) + 1
  uint32_t* z7;        // look-around-7 crossings, bit b of word k <-> idx = 32 k + b
  uint32_t* z3;        // look-around-3 crossings
  int* g_start;        // [cap]
  int* g_len;          // [cap]
  int* count;          // grains found (may exceed cap; only cap rows are written)
};
cudaError_t launch_grain_segment(const GrainSegTrack* tracks_dev, int ntracks, long long max_words, int cap,
                                 cudaStream_t st);

// ---- waveform min/max pyramid (K9)
struct PicksArgs {
  const float* x;          // sample 0 of the padded device copy
  long long n;
  float2* pairs;           // all levels back to back, (min, max) per entry
  long long level_off[32]; // first entry of level l (in pairs); [levels] = total
  int levels;              // levels exist while n > 2^(l+1) (reference app.cpp:352, :365)
};
cudaError_t launch_picks_build(const PicksArgs* tracks_dev, int ntracks, long long max_n, int max_levels,
                               cudaStream_t st);  // one launch for all tracks (blockIdx.y = track)
cudaError_t launch_minmax_ranges(const PicksArgs& a, const int* start_end_dev, int count, float* out_dev,
                                 cudaStream_t st);

}  // namespace mlx
