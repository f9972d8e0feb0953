// melonix_b200/csrc/capi_internal.h -- what the translation units behind the C ABI share: the context
// object, error plumbing and the phase-vocoder launch plan (capi.cu owns the definitions; multi.cu, the
// NCCL layer, uses them).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/melonix_gpu.h"
#include "kernels.h"

namespace mlx {

int fail(int code, const std::string& msg);

#define CK(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return ::mlx::fail(e_ == cudaErrorMemoryAllocation ? MLX_ERR_NOMEM : MLX_ERR_CUDA,  \
                         std::string(#expr) + ": " + cudaGetErrorString(e_));             \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t want) {
    if (want <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) bytes = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

struct Tables {
  DevBuf tw_d, twr_d, tw_f, twr_f, win, win_d, wsyn, decay, tw1_d;
  bool pv_ready = false, spec_ready = false;
};

struct Track {
  size_t offset = 0;  // floats from the base of the track buffer to sample 0
  int64_t n = 0;
};

struct PvPlan {
  int N, H, G, NBP;
  int CA, CS;
  int64_t fb, fe, Fmax;
  int64_t wave_frames;
};

}  // namespace mlx

struct mlx_ctx {
  int device = 0;
  int sm_count = 0, cc = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of mlx_pv_process_host
  int64_t launches = 0;

  // optional per-kernel timing: one event before every launch, one after the last of a sequence
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kind;  // kind of the launch that follows mark i; -1 = end of sequence
  void mark(int kind) {
    if (!profiling) return;
    if (ev_kind.size() == ev_pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      ev_pool.push_back(e);
    }
    cudaEventRecord(ev_pool[ev_kind.size()], stream);
    ev_kind.push_back(kind);
  }

  mlx::DevBuf track_buf;
  std::vector<mlx::Track> tracks;

  std::map<int, mlx::Tables> tables;

  // phase-vocoder scratch
  mlx::DevBuf stage, tot, totc, pre, carry, track_desc, ptr_stage, gk;
  mlx::DevBuf out_wav, out_peak, out_f0;  // device results for the host-pointer entry points
  mlx::DevBuf in_wav16, out_wav16;        // int16 PCM staging of mlx_pv_process_host_fmt
  mlx::DevBuf jobs, spec_out, spec_rgb, spec_desc;
  mlx::DevBuf g_i32a, g_i32b, g_f32a, g_f32b, g_i64, g_out, g_out16;
  mlx::DevBuf seg_bits, seg_desc, seg_rows, seg_count;  // grain segmentation scratch
  mlx::DevBuf picks, picks_ranges, picks_out, picks_desc;  // min/max pyramid of track `picks_track` + query staging
  int picks_track = -1;
  // pinned staging ring for per-call descriptors / tables: a slot is reused only after the copies
  // that read it have completed (event), so launches never wait on the host.
  struct Slot {
    void* p = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
  };
  Slot slots[8];
  int next_slot = 0;

  // analysis left staged by mlx_pv_analyze_dev (K_A output resident in stage / tot / totc)
  struct Staged {
    bool valid = false;
    int N = 0, first = 0, nt = 0, CA = 0;
    float rate = 0.f;
    int64_t fb = 0, fe = 0, wave_frames = 0;
    mlx::PvWave wv{};     // the wave and scratch the analysis ran with (mlx_pv_stage_export_dev)
    mlx::PvScratch sc{};
  } staged;

  const float* track_ptr(int t) const { return static_cast<const float*>(track_buf.p) + tracks[t].offset; }
};

namespace mlx {

// Everything a run needs on the device, staged BEFORE any bulk copy is queued (pv_prepare)
struct PvPrepared {
  const PvTrack* tdev = nullptr;  // [ntracks]
  uint32_t* carry = nullptr;      // [ntracks][NBP]
  Tables* tb = nullptr;
  bool scatter_ok = false;        // the constant-rate pitch-up analysis kernel (K_A2) applies
  bool out16 = false;             // K_S writes PvTrack::out16 (int16 PCM) instead of PvTrack::out
};

enum PvMode { kPvAll = 0, kPvAnalyze = 1, kPvSynth = 2 };
int pv_launch(mlx_ctx* c, const mlx_pv_params* p, const PvPlan& pl, const PvPrepared& pr, int first, int nt,
              bool synth, uint32_t* const* totals_dev, PvMode mode);
int pv_execute(mlx_ctx* c, const mlx_pv_params* p, const PvPlan& pl, bool synth, float* const* out_wav,
               int32_t* const* out_peak, float* const* out_f0, uint32_t* const* totals_dev, PvMode mode);
int acquire_slot(mlx_ctx* c, size_t bytes, mlx_ctx::Slot** out);
int ensure_tables(mlx_ctx* c, int N, bool want_pv, Tables** out);
int layout_tracks(mlx_ctx* c, const int64_t* n, int ntracks);
int64_t num_frames(int64_t n, int hop);
int pv_validate(mlx_ctx* c, const mlx_pv_params* p, PvPlan* pl);
int pv_prepare(mlx_ctx* c, const mlx_pv_params* p, int fftN, bool synth, float* const* out_wav,
               int32_t* const* out_peak, float* const* out_f0, PvPrepared* out, short* const* out_wav16);

}  // namespace mlx
