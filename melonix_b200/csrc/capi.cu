// melonix_b200/csrc/capi.cu -- the extern "C" layer declared in include/melonix_gpu.h.
//
// Owns the device copy of the tracks, twiddle/window tables, scratch for the phase-vocoder
// intermediates and the stream; translates the C calls into kernel launches.  There is no CPU path:
// every entry point either launches sm_100a kernels or returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "capi_internal.h"

using namespace mlx;

namespace mlx {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace mlx

namespace mlx {

// next staging slot of at least `bytes` (waits only if the GPU is 8 calls behind)
int acquire_slot(mlx_ctx* c, size_t bytes, mlx_ctx::Slot** out) {
  mlx_ctx::Slot& s = c->slots[c->next_slot];
  c->next_slot = (c->next_slot + 1) % 8;
  if (!s.done) CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  CK(cudaEventSynchronize(s.done));
  if (s.bytes < bytes) {
    if (s.p) cudaFreeHost(s.p);
    s.p = nullptr;
    s.bytes = 0;
    CK(cudaMallocHost(&s.p, bytes));
    s.bytes = bytes;
  }
  *out = &s;
  return MLX_OK;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <typename T>
int upload_vec(DevBuf& b, const std::vector<T>& v, cudaStream_t st) {
  CK(b.reserve(v.size() * sizeof(T)));
  CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // v is a temporary
  return MLX_OK;
}

// Tables are computed in double on the host (cos/sin at exact angles) and rounded once.
int ensure_tables(mlx_ctx* c, int N, bool want_pv, Tables** out) {
  Tables& tb = c->tables[N];
  const int NC = N / 2;
  if (!tb.spec_ready) {
    std::vector<cplx<float>> twf(NC), twrf(NC / 2 + 1);
    for (int m = 0; m < NC; ++m) {
      const double a = 2.0 * M_PI * (double)m / (double)NC;
      twf[m] = cplx<float>{(float)std::cos(a), (float)-std::sin(a)};
    }
    for (int k = 0; k <= NC / 2; ++k) {
      const double a = 2.0 * M_PI * (double)k / (double)N;
      twrf[k] = cplx<float>{(float)std::cos(a), (float)-std::sin(a)};
    }
    int rc;
    if ((rc = upload_vec(tb.tw_f, twf, c->stream))) return rc;
    if ((rc = upload_vec(tb.twr_f, twrf, c->stream))) return rc;
    // the reference's window factor expf(-2.5e-4f * (start - i)) (spec.cpp:58), indexed by distance
    std::vector<float> decay(N + 1);
    for (int d = 0; d <= N; ++d) decay[d] = expf(-2.5e-4f * (float)d);
    if ((rc = upload_vec(tb.decay, decay, c->stream))) return rc;
    CK(spec_configure(N));
    tb.spec_ready = true;
  }
  if (want_pv && !tb.pv_ready) {
    std::vector<cplx<double>> twd(NC), twrd(NC / 2 + 1);
    for (int m = 0; m < NC; ++m) {
      const double a = 2.0 * M_PI * (double)m / (double)NC;
      twd[m] = cplx<double>{std::cos(a), -std::sin(a)};
    }
    for (int k = 0; k <= NC / 2; ++k) {
      const double a = 2.0 * M_PI * (double)k / (double)N;
      twrd[k] = cplx<double>{std::cos(a), -std::sin(a)};
    }
    // powers of the second FFT stage's twiddle, exp(-2 pi i k r / 256) (fft.cuh: twiddle_table16)
    std::vector<cplx<double>> tw1d(15 * 16);
    for (int r = 1; r < 16; ++r)
      for (int k = 0; k < 16; ++k) {
        const double a = 2.0 * M_PI * (double)(k * r) / 256.0;
        tw1d[(r - 1) * 16 + k] = cplx<double>{std::cos(a), -std::sin(a)};
      }
    // PV-spec A.1: periodic Hann computed in double, stored as float; A.7: g = H / sum w^2 (float)
    std::vector<float> win(N), wsyn(N);
    std::vector<double> wind(N);
    double sw2 = 0.0;
    for (int j = 0; j < N; ++j) {
      win[j] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)j / (double)N));
      wind[j] = (double)win[j];
      sw2 += (double)win[j] * (double)win[j];
    }
    const float g = (float)((double)(N / 4) / sw2);
    for (int j = 0; j < N; ++j) wsyn[j] = (float)((double)g * (double)win[j] / (double)N);
    int rc;
    if ((rc = upload_vec(tb.tw_d, twd, c->stream))) return rc;
    if ((rc = upload_vec(tb.twr_d, twrd, c->stream))) return rc;
    if ((rc = upload_vec(tb.tw1_d, tw1d, c->stream))) return rc;
    if ((rc = upload_vec(tb.win, win, c->stream))) return rc;
    if ((rc = upload_vec(tb.win_d, wind, c->stream))) return rc;
    if ((rc = upload_vec(tb.wsyn, wsyn, c->stream))) return rc;
    CK(pv_configure(N));
    CK(pv_analyze2_configure(N));
    tb.pv_ready = true;
  }
  *out = &tb;
  return MLX_OK;
}

int layout_tracks(mlx_ctx* c, const int64_t* n, int ntracks) {
  if (ntracks <= 0 || !n) return fail(MLX_ERR_INVALID, "ntracks must be > 0");
  c->tracks.assign(ntracks, Track{});
  c->picks_track = -1;  // a cached pyramid belongs to the previous upload
  c->staged.valid = false;
  size_t off = 0;
  for (int t = 0; t < ntracks; ++t) {
    if (n[t] < 0 || n[t] > (int64_t)0x7fffffff - 65536)
      return fail(MLX_ERR_INVALID, "track length must be in [0, 2^31 - 65536) samples (reference indexes with int)");
    c->tracks[t].offset = off + kPadFront;
    c->tracks[t].n = n[t];
    size_t len = (size_t)kPadFront + (size_t)n[t] + (size_t)kPadBack;
    len = (len + 31) & ~size_t(31);  // keep every track 128-byte aligned (TMA needs 16)
    off += len;
  }
  CK(c->track_buf.reserve(off * sizeof(float)));
  CK(cudaMemsetAsync(c->track_buf.p, 0, off * sizeof(float), c->stream));
  return MLX_OK;
}

int64_t num_frames(int64_t n, int hop) { return (n + hop - 1) / hop; }

// K_A2 (csrc/pv_analyze2.cu) is bit-identical to the general analysis kernel but, as measured on a B200
// (profiles/README.md), not yet faster: it runs when MLX_PV_KA2=1 (MLX_PV_NO_KA2=1 always wins).
static bool pv_ka2_enabled() {
  if (getenv("MLX_PV_NO_KA2")) return false;
  const char* e = getenv("MLX_PV_KA2");
  return e && atoi(e) != 0;
}


int pv_validate(mlx_ctx* c, const mlx_pv_params* p, PvPlan* pl) {
  if (!c || !p) return fail(MLX_ERR_INVALID, "null argument");
  if (c->tracks.empty()) return fail(MLX_ERR_STATE, "no tracks uploaded");
  const int N = p->fftN;
  if (!is_pow2(N) || N < 512 || N > kPvMaxN) return fail(MLX_ERR_UNSUPPORTED, "fftN must be 512..8192 (power of two)");
  if (p->hop * 4 != N) return fail(MLX_ERR_UNSUPPORTED, "hop must be fftN/4 (osamp = 4)");
  if (!(p->rate >= 0.25f && p->rate <= 4.0f)) return fail(MLX_ERR_UNSUPPORTED, "rate must be in [0.25, 4]");
  if (!(p->sample_rate > 0)) return fail(MLX_ERR_INVALID, "sample_rate must be > 0");
  pl->N = N;
  pl->H = p->hop;
  pl->G = pv_group_count(N);
  pl->NBP = pv_nbp(N);
  int64_t Fmax = 0;
  for (auto& t : c->tracks) Fmax = std::max(Fmax, num_frames(t.n, p->hop));
  pl->Fmax = Fmax;
  pl->fb = p->frame_begin < 0 ? 0 : p->frame_begin;
  pl->fe = p->frame_end < 0 ? Fmax : std::min<int64_t>(p->frame_end, Fmax);
  if (pl->fb > pl->fe) return fail(MLX_ERR_INVALID, "frame_begin > frame_end");
  // frames per CTA including halo frames: a multiple of G
  int chunk = 256;  // measured: 256 beats 128 and 512 on the 64 x 300 s batch (fewer per-CTA set-ups, same tail)
  if (const char* e = getenv("MLX_PV_CHUNK")) chunk = std::max(8, atoi(e));
  const int64_t span = std::max<int64_t>(1, pl->fe - pl->fb);
  // keep at least ~4 CTAs per SM in flight when the job is small
  while (chunk > 16 && (span / chunk) * (int64_t)c->tracks.size() < 4LL * c->sm_count) chunk /= 2;
  const int GA = pv_group_count_analyze(N);
  pl->CA = GA * std::max(1, chunk / GA) - 1;  // frames per analysis CTA incl. the halo frame: multiple of GA
  int m = std::max(1, chunk / pl->G);
  while (pl->G * m - 3 < 1) ++m;
  pl->CS = pl->G * m - 3;                     // hops per synthesis CTA (+3 halo frames: multiple of G)
  // wave size.  Default: one wave over the whole range (the intermediates stream through HBM, which
  // has >90 % headroom on this compute-bound path) unless that would take more than a quarter of the
  // device memory; an explicit budget tiles the range so that a wave's intermediates stay in L2.
  const double per_frame = (double)c->tracks.size() * pl->NBP * 8.0;
  double budget = -1.0;
  if (p->wave_mib > 0) budget = (double)p->wave_mib * 1048576.0;
  if (p->wave_mib == 0) {
    if (const char* e = getenv("MLX_PV_WAVE_MIB")) {
      if (atoi(e) > 0) budget = (double)atoi(e) * 1048576.0;
    } else if (per_frame * (double)(span + 3) > 0.25 * (double)c->total_mem) {
      budget = 0.25 * (double)c->total_mem;
    }
  }
  if (budget < 0) {
    pl->wave_frames = span;
  } else {
    int64_t wf = (int64_t)(budget / per_frame) - 3;
    wf = std::max<int64_t>(wf, pl->CA);
    pl->wave_frames = std::min<int64_t>(wf, span);
  }
  return MLX_OK;
}

// Stages everything a run needs on the device BEFORE any bulk copy is queued: track descriptors,
// the bin-shift table of the constant rate, zeroed / carried-in phase.  (Small H2D copies issued
// later would queue behind gigabytes of uploads on the single H2D copy engine.)

int pv_prepare(mlx_ctx* c, const mlx_pv_params* p, int fftN, bool synth, float* const* out_wav,
               int32_t* const* out_peak, float* const* out_f0, PvPrepared* out, short* const* out_wav16) {
  const int nt = (int)c->tracks.size();
  const int H = fftN / 4, NBP = pv_nbp(fftN), NC = fftN / 2, NB = NC + 1;
  int rc = ensure_tables(c, fftN, true, &out->tb);
  if (rc) return rc;
  CK(c->track_desc.reserve(sizeof(PvTrack) * nt));
  CK(c->gk.reserve(sizeof(uint32_t) * NBP));
  CK(c->carry.reserve(sizeof(uint32_t) * nt * NBP));
  const size_t desc_bytes = (sizeof(PvTrack) * nt + 15) & ~size_t(15);
  mlx_ctx::Slot* slot = nullptr;
  rc = acquire_slot(c, desc_bytes + sizeof(uint32_t) * NBP, &slot);
  if (rc) return rc;
  PvTrack* desc = static_cast<PvTrack*>(slot->p);
  uint32_t* gk = reinterpret_cast<uint32_t*>(static_cast<char*>(slot->p) + desc_bytes);
  for (int t = 0; t < nt; ++t) {
    desc[t].x = c->track_ptr(t);
    desc[t].n = c->tracks[t].n;
    desc[t].F = num_frames(c->tracks[t].n, H);
    desc[t].out = (synth && out_wav) ? out_wav[t] : nullptr;
    desc[t].out16 = (synth && out_wav16) ? out_wav16[t] : nullptr;
    desc[t].peak = out_peak ? out_peak[t] : nullptr;
    desc[t].f0 = out_f0 ? out_f0[t] : nullptr;
    desc[t].rate_pf = p->rate_per_frame_dev ? p->rate_per_frame_dev[t] : nullptr;
  }
  // bin-shift table for the constant rate (PV-spec A.5): one float multiply per bin, as the spec says
  {
    for (int j = 0; j < NBP; ++j) gk[j] = 1u;  // klo = 1, khi = 0: empty
    std::vector<int> klo(NB, 1), khi(NB, 0);
    const float r = p->rate;
    for (int k = 0; k < NB; ++k) {
      const float tf = (float)k * r;
      const int j = (int)std::trunc(tf);
      if (j < 0 || j >= NB) continue;
      if (klo[j] > khi[j]) klo[j] = k;
      khi[j] = k;
    }
    for (int j = 0; j < NB; ++j)
      if (klo[j] <= khi[j]) gk[j] = (uint32_t)klo[j] | ((uint32_t)khi[j] << 16);
    // K_A2 (opt-in) needs every output bin to be fed by at most one input bin
    bool injective = true;
    for (int j = 0; j < NB; ++j) injective = injective && (klo[j] >= khi[j]);
    if (klo[0] != 0 || khi[0] != 0) injective = false;  // output bin 0 must be fed by input bin 0
    const int kmin = std::max(1, (int)std::ceil(50.0 * fftN / p->sample_rate));
    const int kmax = std::max(kmin, std::min(fftN / 2, (int)std::floor(2000.0 * fftN / p->sample_rate)));
    out->scatter_ok = injective && r >= 1.0f && !p->rate_per_frame_dev && pv_analyze2_supported(fftN) &&
                      kmax - kmin + 1 <= pv_analyze2_band_capacity(fftN) && kmax < fftN / 8 &&
                      pv_ka2_enabled();
  }
  CK(cudaMemcpyAsync(c->track_desc.p, desc, sizeof(PvTrack) * nt, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->gk.p, gk, sizeof(uint32_t) * NBP, cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(slot->done, c->stream));
  CK(cudaMemsetAsync(c->carry.p, 0, sizeof(uint32_t) * nt * NBP, c->stream));
  if (p->phase_in_dev) {
    for (int t = 0; t < nt; ++t)
      if (p->phase_in_dev[t])
        CK(cudaMemcpyAsync(static_cast<uint32_t*>(c->carry.p) + (size_t)t * NBP, p->phase_in_dev[t],
                           sizeof(uint32_t) * NB, cudaMemcpyDeviceToDevice, c->stream));
  }
  out->out16 = synth && out_wav16 != nullptr;
  out->tdev = static_cast<const PvTrack*>(c->track_desc.p);
  out->carry = static_cast<uint32_t*>(c->carry.p);
  return MLX_OK;
}

// Launches waves of K_A -> scan (-> K_S) over frames [fb, fe) of `nt` prepared tracks starting at
// `first`.  Issues kernels only (plus optional D2D copies of the phase totals): nothing here touches
// the H2D copy engine.
//   kPvAll      K_A, scan, K_S per wave (synth = false: no K_S)
//   kPvAnalyze  K_A + scan(carry = 0) over ONE wave; the stage records / tot / totc stay resident in HBM and the
//               carry holds the phase totals of the owned frames
//   kPvSynth    scan(carry = phase_in) + K_S on the staged wave: the scan is re-run because the prefix of
//               every chunk moves with the carried-in phase (it reads tot / totc only: microseconds)
int pv_launch(mlx_ctx* c, const mlx_pv_params* p, const PvPlan& pl, const PvPrepared& pr, int first, int nt,
              bool synth, uint32_t* const* totals_dev, PvMode mode) {
  Tables* tb = pr.tb;
  const size_t rows = (size_t)pl.wave_frames + 3;
  const size_t nchunksA_max = (size_t)((pl.wave_frames + 3 + pl.CA - 1) / pl.CA);
  if (mode != kPvAll && pl.wave_frames < pl.fe - pl.fb)
    return fail(MLX_ERR_UNSUPPORTED, "the split analyze / synth calls need the frame range in one wave (wave_mib < 0)");
  if (mode == kPvSynth) {
    const mlx_ctx::Staged& sg = c->staged;
    if (!sg.valid || sg.N != pl.N || sg.rate != p->rate || sg.fb != pl.fb || sg.fe != pl.fe || sg.first != first ||
        sg.nt != nt || sg.CA != pl.CA || sg.wave_frames != pl.wave_frames)
      return fail(MLX_ERR_STATE, "mlx_pv_synth_dev: no staged analysis with these parameters (call mlx_pv_analyze_dev first)");
  } else {
    c->staged.valid = false;  // the scratch is about to be overwritten
    CK(c->stage.reserve(sizeof(uint2) * nt * rows * pl.NBP));
    CK(c->tot.reserve(sizeof(uint32_t) * nt * nchunksA_max * pl.NBP));
    CK(c->totc.reserve(sizeof(uint32_t) * nt * nchunksA_max * pl.NBP));
    CK(c->pre.reserve(sizeof(uint32_t) * nt * nchunksA_max * pl.NBP));
  }

  PvTables pt{static_cast<const cplx<double>*>(tb->tw_d.p), static_cast<const cplx<double>*>(tb->twr_d.p),
              static_cast<const cplx<float>*>(tb->tw_f.p),  static_cast<const cplx<float>*>(tb->twr_f.p),
              static_cast<const float*>(tb->win.p),         static_cast<const double*>(tb->win_d.p),
              static_cast<const float*>(tb->wsyn.p),        static_cast<const cplx<double>*>(tb->tw1_d.p)};
  PvScratch sc{static_cast<uint2*>(c->stage.p), static_cast<uint32_t*>(c->tot.p),
               static_cast<uint32_t*>(c->totc.p), static_cast<uint32_t*>(c->pre.p),
               pr.carry + (size_t)first * pl.NBP};
  const PvTrack* tdev = pr.tdev + first;

  int kmin = (int)std::ceil(50.0 * pl.N / p->sample_rate), kmax = (int)std::floor(2000.0 * pl.N / p->sample_rate);
  kmin = std::max(kmin, 1);
  kmax = std::min(kmax, pl.N / 2);
  kmax = std::max(kmax, kmin);

  PvWave last_wv{};
  for (int64_t wb = pl.fb; wb < pl.fe; wb += pl.wave_frames) {
    PvWave wv{};
    wv.wb = wb;
    wv.we = std::min<int64_t>(wb + pl.wave_frames, pl.fe);
    wv.rows = (int)rows;
    wv.CA = pl.CA;
    wv.nchunksA = (int)((wv.we + 3 - wv.wb + pl.CA - 1) / pl.CA);
    wv.CS = pl.CS;
    wv.nchunksS = (int)((wv.we - wv.wb + pl.CS - 1) / pl.CS);
    wv.rate = p->rate;
    wv.fs_over_N = (float)(p->sample_rate / (double)pl.N);
    wv.kmin = kmin;
    wv.kmax = kmax;
    wv.gk = static_cast<const uint32_t*>(c->gk.p);
    wv.r_fix = (long long)((double)p->rate * 67108864.0);
    if (mode != kPvSynth) {
      c->mark(0);
      if (pr.scatter_ok)
        CK(launch_pv_analyze2(pl.N, tdev, nt, wv, pt, sc, c->stream));  // constant ratio >= 1: K_A2 (opt-in)
      else
        CK(launch_pv_analyze(pl.N, tdev, nt, wv, pt, sc, !p->rate_per_frame_dev && p->rate >= 1.0f, c->stream));
      c->launches += 1;
    }
    c->mark(1);
    CK(launch_pv_scan(pl.N, nt, wv, sc, c->stream));
    c->launches += 1;
    last_wv = wv;
    if (synth && mode != kPvAnalyze) {
      c->mark(2);
      CK(launch_pv_synth(pl.N, tdev, nt, wv, pt, sc, pr.out16, c->stream));
      c->launches += 1;
    }
  }
  c->mark(-1);
  if (mode == kPvAnalyze) {
    mlx_ctx::Staged& sg = c->staged;
    sg.valid = true;
    sg.N = pl.N;
    sg.rate = p->rate;
    sg.fb = pl.fb;
    sg.fe = pl.fe;
    sg.first = first;
    sg.nt = nt;
    sg.CA = pl.CA;
    sg.wave_frames = pl.wave_frames;
    sg.wv = last_wv;
    sg.sc = sc;
  }
  if (totals_dev) {
    for (int t = 0; t < nt; ++t)
      if (totals_dev[t])
        CK(cudaMemcpyAsync(totals_dev[t], sc.carry + (size_t)t * pl.NBP, sizeof(uint32_t) * (pl.N / 2 + 1),
                           cudaMemcpyDeviceToDevice, c->stream));
  }
  return MLX_OK;
}

// prepare + launch over all uploaded tracks
int pv_execute(mlx_ctx* c, const mlx_pv_params* p, const PvPlan& pl, bool synth, float* const* out_wav,
               int32_t* const* out_peak, float* const* out_f0, uint32_t* const* totals_dev, PvMode mode) {
  PvPrepared pr;
  int rc = pv_prepare(c, p, pl.N, synth, out_wav, out_peak, out_f0, &pr, nullptr);
  if (rc) return rc;
  return pv_launch(c, p, pl, pr, 0, (int)c->tracks.size(), synth, totals_dev, mode);
}

}  // namespace mlx

extern "C" {

const char* mlx_last_error(void) { return mlx::g_err.c_str(); }

int mlx_create(mlx_ctx** out, int device) {
  if (!out) return fail(MLX_ERR_INVALID, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(MLX_ERR_CUDA, std::string("no CUDA device: melonix_b200 has no CPU fallback (") +
                                  cudaGetErrorString(e) + ")");
  if (device < 0 || device >= ndev) return fail(MLX_ERR_INVALID, "device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MLX_ERR_UNSUPPORTED, std::string("kernels are built for sm_100a only; device is ") + prop.name);
  mlx_ctx* c = new mlx_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc = prop.major * 10 + prop.minor;
  c->total_mem = prop.totalGlobalMem;
  if (cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return fail(MLX_ERR_CUDA, "cudaStreamCreate failed");
  }
  *out = c;
  return MLX_OK;
}

void mlx_destroy(mlx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (DevBuf* b : {&c->gk, &c->track_buf, &c->stage, &c->tot, &c->totc, &c->pre, &c->carry, &c->track_desc, &c->ptr_stage,
                    &c->out_wav, &c->out_peak, &c->out_f0, &c->in_wav16, &c->out_wav16, &c->spec_desc, &c->jobs, &c->spec_out, &c->spec_rgb, &c->g_i32a,
                    &c->g_i32b, &c->g_f32a, &c->g_f32b, &c->g_i64, &c->g_out, &c->g_out16, &c->seg_bits,
                    &c->seg_desc, &c->seg_rows, &c->seg_count, &c->picks, &c->picks_ranges, &c->picks_out, &c->picks_desc})
    b->release();
  for (auto& kv : c->tables)
    for (DevBuf* b : {&kv.second.tw_d, &kv.second.twr_d, &kv.second.tw_f, &kv.second.twr_f, &kv.second.win,
                      &kv.second.win_d, &kv.second.wsyn, &kv.second.decay, &kv.second.tw1_d})
      b->release();
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  for (auto& s : c->slots) {
    if (s.p) cudaFreeHost(s.p);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
}

int mlx_set_stream(mlx_ctx* c, void* cuda_stream) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  return MLX_OK;
}

int mlx_sync(mlx_ctx* c) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

int mlx_device_info(mlx_ctx* c, int* sm_count, int* cc, size_t* total_mem) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  if (sm_count) *sm_count = c->sm_count;
  if (cc) *cc = c->cc;
  if (total_mem) *total_mem = c->total_mem;
  return MLX_OK;
}

int64_t mlx_launch_count(const mlx_ctx* c) { return c ? c->launches : 0; }

int mlx_profile_enable(mlx_ctx* c, int on) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  c->profiling = on != 0;
  return MLX_OK;
}

int mlx_profile_read(mlx_ctx* c, double* ms, int64_t* launches, int reset) {
  if (!c || !ms || !launches) return fail(MLX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 5; ++k) {
    ms[k] = 0.0;
    launches[k] = 0;
  }
  for (size_t i = 0; i + 1 < c->ev_kind.size(); ++i) {
    const int kind = c->ev_kind[i];
    if (kind < 0 || kind >= 5) continue;
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, c->ev_pool[i], c->ev_pool[i + 1]));
    ms[kind] += t;
    launches[kind] += 1;
  }
  if (reset) c->ev_kind.clear();
  return MLX_OK;
}

int mlx_upload_tracks(mlx_ctx* c, const float* const* wav, const int64_t* n, int ntracks) {
  if (!c || !wav) return fail(MLX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  int rc = layout_tracks(c, n, ntracks);
  if (rc) return rc;
  for (int t = 0; t < ntracks; ++t)
    if (n[t] > 0)
      CK(cudaMemcpyAsync(static_cast<float*>(c->track_buf.p) + c->tracks[t].offset, wav[t], sizeof(float) * n[t],
                         cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

int mlx_upload_tracks_dev(mlx_ctx* c, const float* const* wav_dev, const int64_t* n, int ntracks) {
  if (!c || !wav_dev) return fail(MLX_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  int rc = layout_tracks(c, n, ntracks);
  if (rc) return rc;
  for (int t = 0; t < ntracks; ++t)
    if (n[t] > 0)
      CK(cudaMemcpyAsync(static_cast<float*>(c->track_buf.p) + c->tracks[t].offset, wav_dev[t],
                         sizeof(float) * n[t], cudaMemcpyDeviceToDevice, c->stream));
  return MLX_OK;
}

int mlx_num_tracks(const mlx_ctx* c) { return c ? (int)c->tracks.size() : 0; }
int64_t mlx_track_len(const mlx_ctx* c, int t) {
  return (c && t >= 0 && t < (int)c->tracks.size()) ? c->tracks[t].n : -1;
}

// ------------------------------------------------------------------------------------------------ Spec
static int spec_common(mlx_ctx* c, int track, int fftN, const int* jobs_dev, int hop, int64_t first_frame,
                       int64_t count, float* out_dev, unsigned char* rgb_dev, float k) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  if (track < 0 || track >= (int)c->tracks.size()) return fail(MLX_ERR_STATE, "track not uploaded");
  if (!is_pow2(fftN) || fftN < 512 || fftN > 32768)
    return fail(MLX_ERR_UNSUPPORTED, "fftN must be a power of two in [512, 32768]");
  if (count < 0) return fail(MLX_ERR_INVALID, "count < 0");
  CK(cudaSetDevice(c->device));
  Tables* tb = nullptr;
  int rc = ensure_tables(c, fftN, false, &tb);
  if (rc) return rc;
  SpecArgs a{};
  a.x = c->track_ptr(track);
  a.n = c->tracks[track].n;
  a.jobs = jobs_dev;
  a.hop = hop;
  a.first_frame = first_frame;
  a.count = count;
  a.out = out_dev;
  a.rgb = rgb_dev;
  a.kcol = k;
  a.tw_f = static_cast<const cplx<float>*>(tb->tw_f.p);
  a.twr_f = static_cast<const cplx<float>*>(tb->twr_f.p);
  a.decay = static_cast<const float*>(tb->decay.p);
  c->mark(3);
  CK(launch_spec(fftN, a, c->stream));
  c->mark(-1);
  if (count > 0) c->launches += 1;
  return MLX_OK;
}

int mlx_spec_batch_dev(mlx_ctx* c, int track, int fftN, const int32_t* start_end_dev, int count, float* out_dev) {
  if (!start_end_dev || !out_dev) return fail(MLX_ERR_INVALID, "null argument");
  return spec_common(c, track, fftN, start_end_dev, 0, 0, count, out_dev, nullptr, 0.f);
}

int mlx_spec_frames_dev(mlx_ctx* c, int track, int fftN, int hop, int64_t first_frame, int64_t count,
                        float* out_dev) {
  if (!out_dev || hop <= 0) return fail(MLX_ERR_INVALID, "bad argument");
  return spec_common(c, track, fftN, nullptr, hop, first_frame, count, out_dev, nullptr, 0.f);
}

// A host job list that is a run of consecutive frames of one hop -- jobs[i] = (s0 + i*hop,
// s0 + (i+1)*hop) with s0 a multiple of hop, which is what SpecCache::populateTex produces at a fixed
// zoom (spec-cache.cpp:63-65) -- is handed to the kernels in regular-hop form, so that launch_spec can
// pick the tiled kernel; any other list stays a list.
int mlx_spec_frames_all_dev(mlx_ctx* c, int fftN, int hop, float* const* out_dev) {
  if (!c || !out_dev) return fail(MLX_ERR_INVALID, "null argument");
  if (c->tracks.empty()) return fail(MLX_ERR_STATE, "no tracks uploaded");
  if (hop <= 0) return fail(MLX_ERR_INVALID, "hop must be > 0");
  const int nt = (int)c->tracks.size();
  // the batched launch needs the TMA-tiled regular-hop kernel (fftN <= 8192, hop a multiple of 4 and
  // <= fftN); anything else goes track by track through the general path
  if (fftN > 8192 || (hop & 3) != 0 || hop > fftN || getenv("MLX_SPEC_GENERIC")) {
    for (int t = 0; t < nt; ++t) {
      int rc = spec_common(c, t, fftN, nullptr, hop, 0, num_frames(c->tracks[t].n, hop), out_dev[t], nullptr, 0.f);
      if (rc) return rc;
    }
    return MLX_OK;
  }
  if (!is_pow2(fftN) || fftN < 512) return fail(MLX_ERR_UNSUPPORTED, "fftN must be a power of two in [512, 32768]");
  CK(cudaSetDevice(c->device));
  Tables* tb = nullptr;
  int rc = ensure_tables(c, fftN, false, &tb);
  if (rc) return rc;
  CK(c->spec_desc.reserve(sizeof(SpecTrackDesc) * nt));
  mlx_ctx::Slot* slot = nullptr;
  rc = acquire_slot(c, sizeof(SpecTrackDesc) * nt, &slot);
  if (rc) return rc;
  SpecTrackDesc* d = static_cast<SpecTrackDesc*>(slot->p);
  int64_t fmax = 0;
  int tmax = 0;
  for (int t = 0; t < nt; ++t) {
    d[t].x = c->track_ptr(t);
    d[t].n = c->tracks[t].n;
    d[t].count = num_frames(c->tracks[t].n, hop);
    d[t].out = out_dev[t];
    d[t].rgb = nullptr;
    if (d[t].count > fmax) {
      fmax = d[t].count;
      tmax = t;
    }
  }
  CK(cudaMemcpyAsync(c->spec_desc.p, d, sizeof(SpecTrackDesc) * nt, cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(slot->done, c->stream));
  SpecArgs a{};
  a.x = c->track_ptr(tmax);  // (placeholders of the longest track: the launch geometry is sized for it)
  a.n = c->tracks[tmax].n;
  a.hop = hop;
  a.first_frame = 0;
  a.count = fmax;
  a.out = out_dev[tmax];
  a.tw_f = static_cast<const cplx<float>*>(tb->tw_f.p);
  a.twr_f = static_cast<const cplx<float>*>(tb->twr_f.p);
  a.decay = static_cast<const float*>(tb->decay.p);
  a.multi = static_cast<const SpecTrackDesc*>(c->spec_desc.p);
  a.ntracks = nt;
  c->mark(3);
  CK(launch_spec(fftN, a, c->stream));
  c->mark(-1);
  if (fmax > 0) c->launches += 1;
  return MLX_OK;
}

static bool regular_run(const int32_t* se, int count, int* hop, int64_t* first_frame) {
  const int64_t h = (int64_t)se[1] - se[0];
  if (h <= 0 || se[0] < 0 || se[0] % h != 0) return false;
  for (int i = 0; i < count; ++i)
    if (se[2 * i] != se[0] + i * h || se[2 * i + 1] != se[2 * i] + h) return false;
  *hop = (int)h;
  *first_frame = se[0] / h;
  return true;
}

int mlx_spec_batch(mlx_ctx* c, int track, int fftN, const int32_t* start_end, int count, float* out) {
  if (!c || !start_end || !out) return fail(MLX_ERR_INVALID, "null argument");
  if (count <= 0) return count == 0 ? MLX_OK : fail(MLX_ERR_INVALID, "count < 0");
  CK(cudaSetDevice(c->device));
  CK(c->jobs.reserve(sizeof(int32_t) * 2 * (size_t)count));
  CK(c->spec_out.reserve(sizeof(float) * (size_t)count * (fftN / 2)));
  int hop = 0, rc;
  int64_t first = 0;
  if (regular_run(start_end, count, &hop, &first)) {
    rc = spec_common(c, track, fftN, nullptr, hop, first, count, static_cast<float*>(c->spec_out.p), nullptr, 0.f);
  } else {
    CK(cudaMemcpyAsync(c->jobs.p, start_end, sizeof(int32_t) * 2 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    rc = spec_common(c, track, fftN, static_cast<const int*>(c->jobs.p), 0, 0, count,
                     static_cast<float*>(c->spec_out.p), nullptr, 0.f);
  }
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, c->spec_out.p, sizeof(float) * (size_t)count * (fftN / 2), cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

int mlx_spec_batch_rgb(mlx_ctx* c, int track, int fftN, const int32_t* start_end, int count, float k,
                       uint8_t* out_rgb) {
  if (!c || !start_end || !out_rgb) return fail(MLX_ERR_INVALID, "null argument");
  if (count <= 0) return count == 0 ? MLX_OK : fail(MLX_ERR_INVALID, "count < 0");
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)count * (fftN / 2) * 3;
  CK(c->jobs.reserve(sizeof(int32_t) * 2 * (size_t)count));
  CK(c->spec_rgb.reserve(bytes));
  int hop = 0, rc;
  int64_t first = 0;
  if (regular_run(start_end, count, &hop, &first)) {
    rc = spec_common(c, track, fftN, nullptr, hop, first, count, nullptr, static_cast<unsigned char*>(c->spec_rgb.p), k);
  } else {
    CK(cudaMemcpyAsync(c->jobs.p, start_end, sizeof(int32_t) * 2 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    rc = spec_common(c, track, fftN, static_cast<const int*>(c->jobs.p), 0, 0, count, nullptr,
                     static_cast<unsigned char*>(c->spec_rgb.p), k);
  }
  if (rc) return rc;
  CK(cudaMemcpyAsync(out_rgb, c->spec_rgb.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

// ------------------------------------------------------------------------------------------------ PV
int mlx_pv_run_dev(mlx_ctx* c, const mlx_pv_params* p, float* const* out_wav_dev, int32_t* const* out_peak_dev,
                   float* const* out_f0_dev) {
  PvPlan pl{};
  int rc = pv_validate(c, p, &pl);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  return pv_execute(c, p, pl, true, out_wav_dev, out_peak_dev, out_f0_dev, nullptr, kPvAll);
}

int mlx_pv_phase_totals_dev(mlx_ctx* c, const mlx_pv_params* p, uint32_t* const* totals_dev,
                            int32_t* const* out_peak_dev, float* const* out_f0_dev) {
  PvPlan pl{};
  int rc = pv_validate(c, p, &pl);
  if (rc) return rc;
  if (!totals_dev) return fail(MLX_ERR_INVALID, "totals_dev is null");
  CK(cudaSetDevice(c->device));
  mlx_pv_params q = *p;
  q.phase_in_dev = nullptr;  // totals are relative to the first owned frame
  return pv_execute(c, &q, pl, false, nullptr, out_peak_dev, out_f0_dev, totals_dev, kPvAll);
}

int mlx_pv_analyze_dev(mlx_ctx* c, const mlx_pv_params* p, uint32_t* const* totals_dev,
                       int32_t* const* out_peak_dev, float* const* out_f0_dev) {
  PvPlan pl{};
  mlx_pv_params q{};
  if (p) {
    q = *p;
    q.wave_mib = -1;  // the staged intermediates must cover the whole range
    q.phase_in_dev = nullptr;
  }
  int rc = pv_validate(c, p ? &q : nullptr, &pl);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  return pv_execute(c, &q, pl, false, nullptr, out_peak_dev, out_f0_dev, totals_dev, kPvAnalyze);
}

int mlx_pv_synth_dev(mlx_ctx* c, const mlx_pv_params* p, float* const* out_wav_dev) {
  PvPlan pl{};
  mlx_pv_params q{};
  if (p) {
    q = *p;
    q.wave_mib = -1;
  }
  int rc = pv_validate(c, p ? &q : nullptr, &pl);
  if (rc) return rc;
  if (!out_wav_dev) return fail(MLX_ERR_INVALID, "out_wav_dev is null");
  CK(cudaSetDevice(c->device));
  return pv_execute(c, &q, pl, true, out_wav_dev, nullptr, nullptr, nullptr, kPvSynth);
}

int mlx_pv_stage_export_dev(mlx_ctx* c, int track, int64_t frame_begin, int64_t count, float* smag_dev,
                            uint32_t* phase_dev) {
  if (!c || !smag_dev || !phase_dev) return fail(MLX_ERR_INVALID, "null argument");
  const mlx_ctx::Staged& sg = c->staged;
  if (!sg.valid) return fail(MLX_ERR_STATE, "mlx_pv_stage_export_dev: no staged analysis (call mlx_pv_analyze_dev first)");
  if (track < sg.first || track >= sg.first + sg.nt) return fail(MLX_ERR_INVALID, "track is not part of the staged analysis");
  if (count < 0 || frame_begin < sg.fb || frame_begin + count > sg.fe)
    return fail(MLX_ERR_INVALID, "frames outside the staged range");
  CK(cudaSetDevice(c->device));
  CK(launch_pv_stage_export(sg.N, track - sg.first, sg.wv, sg.sc, frame_begin, count, smag_dev, phase_dev, c->stream));
  c->launches += count > 0;
  return MLX_OK;
}

int mlx_pv_run(mlx_ctx* c, const mlx_pv_params* p, float* const* out_wav, int32_t* const* out_peak,
               float* const* out_f0) {
  PvPlan pl{};
  int rc = pv_validate(c, p, &pl);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const int nt = (int)c->tracks.size();
  // device result buffers, one slab per kind
  std::vector<size_t> woff(nt + 1, 0), foff(nt + 1, 0);
  for (int t = 0; t < nt; ++t) {
    woff[t + 1] = woff[t] + (((size_t)c->tracks[t].n + 31) & ~size_t(31));
    foff[t + 1] = foff[t] + (((size_t)num_frames(c->tracks[t].n, pl.H) + 31) & ~size_t(31));
  }
  if (out_wav) CK(c->out_wav.reserve(sizeof(float) * std::max<size_t>(woff[nt], 1)));
  if (out_peak) CK(c->out_peak.reserve(sizeof(int32_t) * std::max<size_t>(foff[nt], 1)));
  if (out_f0) CK(c->out_f0.reserve(sizeof(float) * std::max<size_t>(foff[nt], 1)));
  std::vector<float*> dw(nt, nullptr), df(nt, nullptr);
  std::vector<int32_t*> dp(nt, nullptr);
  for (int t = 0; t < nt; ++t) {
    if (out_wav && out_wav[t]) dw[t] = static_cast<float*>(c->out_wav.p) + woff[t];
    if (out_peak && out_peak[t]) dp[t] = static_cast<int32_t*>(c->out_peak.p) + foff[t];
    if (out_f0 && out_f0[t]) df[t] = static_cast<float*>(c->out_f0.p) + foff[t];
  }
  rc = pv_execute(c, p, pl, true, out_wav ? dw.data() : nullptr, out_peak ? dp.data() : nullptr,
                  out_f0 ? df.data() : nullptr, nullptr, kPvAll);
  if (rc) return rc;
  // copy back only what this call owns: hops / frames [fb, fe)
  for (int t = 0; t < nt; ++t) {
    const int64_t F = num_frames(c->tracks[t].n, pl.H);
    const int64_t f0 = std::min(pl.fb, F), f1 = std::min(pl.fe, F);
    if (f1 <= f0) continue;
    const int64_t s0 = f0 * pl.H, s1 = std::min<int64_t>(f1 * pl.H, c->tracks[t].n);
    if (dw[t] && s1 > s0)
      CK(cudaMemcpyAsync(out_wav[t] + s0, dw[t] + s0, sizeof(float) * (s1 - s0), cudaMemcpyDeviceToHost, c->stream));
    if (dp[t])
      CK(cudaMemcpyAsync(out_peak[t] + f0, dp[t] + f0, sizeof(int32_t) * (f1 - f0), cudaMemcpyDeviceToHost,
                         c->stream));
    if (df[t])
      CK(cudaMemcpyAsync(out_f0[t] + f0, df[t] + f0, sizeof(float) * (f1 - f0), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

// Host buffers in, host buffers out, copies overlapped with the kernels: uploads on s_in, kernels on the
// context stream, downloads on s_out.  Tracks go through the pipeline in groups (one K_A / scan / K_S
// launch sequence per group: enough CTAs per launch to fill the chip, few enough tracks per group that the
// first download starts early).  Sample formats: float32, or int16 PCM on either side -- x = s / 32768 in
// (what the reference's decoder produces, swr s16 -> flt), int16(x * 32767.) out (the reference's export
// conversion, app.cpp:1209-1212, fused into K_S) -- which halves the bytes on the PCIe wire.
static int pv_process_host_impl(mlx_ctx* c, const mlx_pv_params* p, const void* const* wav, int in_fmt,
                                const int64_t* n, int ntracks, void* const* out_wav, int out_fmt,
                                int32_t* const* out_peak, float* const* out_f0) {
  if (!c || !p || !wav || !n || ntracks <= 0) return fail(MLX_ERR_INVALID, "bad argument");
  if ((in_fmt != MLX_FMT_F32 && in_fmt != MLX_FMT_I16) || (out_fmt != MLX_FMT_F32 && out_fmt != MLX_FMT_I16))
    return fail(MLX_ERR_INVALID, "sample format must be MLX_FMT_F32 or MLX_FMT_I16");
  if (p->frame_begin > 0 || p->frame_end >= 0 || p->phase_in_dev || p->rate_per_frame_dev)
    return fail(MLX_ERR_UNSUPPORTED, "mlx_pv_process_host handles whole tracks with a constant rate");
  CK(cudaSetDevice(c->device));
  int rc = layout_tracks(c, n, ntracks);
  if (rc) return rc;
  const std::vector<Track> all = c->tracks;
  const int H = p->hop;
  const bool in16 = in_fmt == MLX_FMT_I16, out16 = out_fmt == MLX_FMT_I16;
  std::vector<size_t> woff(ntracks + 1, 0), foff(ntracks + 1, 0);
  for (int t = 0; t < ntracks; ++t) {
    woff[t + 1] = woff[t] + (((size_t)n[t] + 31) & ~size_t(31));
    foff[t + 1] = foff[t] + (((size_t)num_frames(n[t], H) + 31) & ~size_t(31));
  }
  if (out16)
    CK(c->out_wav16.reserve(sizeof(short) * std::max<size_t>(woff[ntracks], 1)));
  else
    CK(c->out_wav.reserve(sizeof(float) * std::max<size_t>(woff[ntracks], 1)));
  if (in16) CK(c->in_wav16.reserve(sizeof(short) * std::max<size_t>(woff[ntracks], 1)));
  CK(c->out_peak.reserve(sizeof(int32_t) * std::max<size_t>(foff[ntracks], 1)));
  CK(c->out_f0.reserve(sizeof(float) * std::max<size_t>(foff[ntracks], 1)));
  std::vector<float*> dw(ntracks, nullptr), df(ntracks, nullptr);
  std::vector<short*> dw16(ntracks, nullptr);
  std::vector<int32_t*> dp(ntracks, nullptr);
  for (int t = 0; t < ntracks; ++t) {
    if (out_wav && out_wav[t]) {
      if (out16)
        dw16[t] = static_cast<short*>(c->out_wav16.p) + woff[t];
      else
        dw[t] = static_cast<float*>(c->out_wav.p) + woff[t];
    }
    if (out_peak && out_peak[t]) dp[t] = static_cast<int32_t*>(c->out_peak.p) + foff[t];
    if (out_f0 && out_f0[t]) df[t] = static_cast<float*>(c->out_f0.p) + foff[t];
  }
  // validate once (sizes, ratio) and stage descriptors / table / phase for ALL tracks up front
  PvPlan pl_all{};
  rc = pv_validate(c, p, &pl_all);
  if (rc) return rc;
  PvPrepared pr;
  rc = pv_prepare(c, p, p->fftN, true, out16 ? nullptr : dw.data(), dp.data(), df.data(), &pr,
                  out16 ? dw16.data() : nullptr);
  if (rc) return rc;
  // tracks per launch group.  Measured on the bench batch with int16 on the wire (tools/e2e_probe.py: the
  // pure-copy floor of the same bytes is 40.5 ms): 1 / 2 / 4 / 8 / 16 tracks per group -> 43.4 / 43.9 / 45.2 /
  // 47.3 / 53.0 ms.  Small groups win -- the kernels stay hidden under the copies either way and the drain
  // after the last upload is shorter; 2 keeps some margin for the kernels (222 CTAs per launch).
  int grp = 2;
  if (const char* e = getenv("MLX_PV_HOST_GROUP")) grp = std::max(1, atoi(e));
  grp = std::min(grp, ntracks);
  // scratch for the largest group, allocated before the pipeline starts (no cudaMalloc between launches)
  {
    size_t rows_nbp = 0, nch_nbp = 0;
    for (int t0 = 0; t0 < ntracks; t0 += grp) {
      const int g = std::min(grp, ntracks - t0);
      c->tracks.assign(all.begin() + t0, all.begin() + t0 + g);
      PvPlan plg{};
      rc = pv_validate(c, p, &plg);
      if (rc) break;
      const size_t rows = (size_t)plg.wave_frames + 3, nch = (size_t)((plg.wave_frames + 3 + plg.CA - 1) / plg.CA);
      rows_nbp = std::max(rows_nbp, rows * plg.NBP * g);
      nch_nbp = std::max(nch_nbp, nch * plg.NBP * g);
    }
    c->tracks = all;
    if (rc) return rc;
    CK(c->stage.reserve(sizeof(uint2) * rows_nbp));
    CK(c->tot.reserve(sizeof(uint32_t) * nch_nbp));
    CK(c->totc.reserve(sizeof(uint32_t) * nch_nbp));
    CK(c->pre.reserve(sizeof(uint32_t) * nch_nbp));
  }
  const int ngroups = (ntracks + grp - 1) / grp;
  std::vector<cudaEvent_t> ev_in(ngroups), ev_done(ngroups);
  for (int g = 0; g < ngroups; ++g) {
    cudaEventCreateWithFlags(&ev_in[g], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_done[g], cudaEventDisableTiming);
  }
  cudaEvent_t ev_zero;
  cudaEventCreate(&ev_zero);
  cudaEventRecord(ev_zero, c->stream);  // padding memset + staging copies precede the uploads
  cudaStreamWaitEvent(c->s_in, ev_zero, 0);
  const bool trace = getenv("MLX_TRACE") != nullptr;  // development aid: pipeline timeline on stderr
  cudaEvent_t tr_in_end = nullptr, tr_k_first = nullptr, tr_k_last = nullptr, tr_out_first = nullptr, tr_out_last = nullptr;
  if (trace)
    for (cudaEvent_t* e : {&tr_in_end, &tr_k_first, &tr_k_last, &tr_out_first, &tr_out_last}) cudaEventCreate(e);
  int result = MLX_OK;
  // all uploads are queued first (the copy engine runs them back to back) ...
  for (int t = 0; t < ntracks && result == MLX_OK; ++t) {
    if (n[t] > 0) {
      cudaError_t e;
      if (in16)
        e = cudaMemcpyAsync(static_cast<short*>(c->in_wav16.p) + woff[t], wav[t], sizeof(short) * n[t],
                            cudaMemcpyHostToDevice, c->s_in);
      else
        e = cudaMemcpyAsync(static_cast<float*>(c->track_buf.p) + all[t].offset, wav[t], sizeof(float) * n[t],
                            cudaMemcpyHostToDevice, c->s_in);
      if (e != cudaSuccess) result = fail(MLX_ERR_CUDA, "H2D copy failed");
    }
    if ((t + 1) % grp == 0 || t + 1 == ntracks) cudaEventRecord(ev_in[t / grp], c->s_in);
  }
  if (trace) cudaEventRecord(tr_in_end, c->s_in);
  // ... then, group by group: wait for its samples, (convert,) analyse + synthesise, download
  for (int g = 0; g < ngroups && result == MLX_OK; ++g) {
    const int t0 = g * grp, gn = std::min(grp, ntracks - t0);
    cudaStreamWaitEvent(c->stream, ev_in[g], 0);
    if (in16) {
      for (int t = t0; t < t0 + gn && result == MLX_OK; ++t) {
        if (launch_pcm16_to_float(static_cast<const short*>(c->in_wav16.p) + woff[t],
                                  static_cast<float*>(c->track_buf.p) + all[t].offset, n[t], c->stream) != cudaSuccess)
          result = fail(MLX_ERR_CUDA, "pcm16_to_float launch failed");
        c->launches += n[t] > 0;
      }
      if (result) break;
    }
    c->tracks.assign(all.begin() + t0, all.begin() + t0 + gn);  // plan (chunk sizes) for this group; kernels only from here on
    PvPlan pl{};
    result = pv_validate(c, p, &pl);
    if (result) break;
    result = pv_launch(c, p, pl, pr, t0, gn, true, nullptr, kPvAll);
    if (result) break;
    cudaEventRecord(ev_done[g], c->stream);
    if (trace && g == 0) cudaEventRecord(tr_k_first, c->stream);
    if (trace && g == ngroups - 1) cudaEventRecord(tr_k_last, c->stream);
    cudaStreamWaitEvent(c->s_out, ev_done[g], 0);
    for (int t = t0; t < t0 + gn; ++t) {
      const int64_t F = num_frames(n[t], H);
      if (dw[t] && n[t] > 0) cudaMemcpyAsync(out_wav[t], dw[t], sizeof(float) * n[t], cudaMemcpyDeviceToHost, c->s_out);
      if (dw16[t] && n[t] > 0) cudaMemcpyAsync(out_wav[t], dw16[t], sizeof(short) * n[t], cudaMemcpyDeviceToHost, c->s_out);
      if (dp[t] && F > 0) cudaMemcpyAsync(out_peak[t], dp[t], sizeof(int32_t) * F, cudaMemcpyDeviceToHost, c->s_out);
      if (df[t] && F > 0) cudaMemcpyAsync(out_f0[t], df[t], sizeof(float) * F, cudaMemcpyDeviceToHost, c->s_out);
    }
    if (trace && g == 0) cudaEventRecord(tr_out_first, c->s_out);
    if (trace && g == ngroups - 1) cudaEventRecord(tr_out_last, c->s_out);
  }
  c->tracks = all;
  c->staged.valid = false;
  cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->s_out),
              e3 = cudaStreamSynchronize(c->s_in);
  if (trace && result == MLX_OK) {
    float a = 0, b = 0, d = 0, e = 0, f = 0;
    cudaEventElapsedTime(&a, ev_zero, tr_in_end);
    cudaEventElapsedTime(&b, ev_zero, tr_k_first);
    cudaEventElapsedTime(&d, ev_zero, tr_k_last);
    cudaEventElapsedTime(&e, ev_zero, tr_out_first);
    cudaEventElapsedTime(&f, ev_zero, tr_out_last);
    fprintf(stderr, "[mlx trace] ms since start: last H2D done %.1f | kernels of group 0 done %.1f, last group %.1f | "
                    "D2H of group 0 done %.1f, last %.1f\n", a, b, d, e, f);
    for (cudaEvent_t ev : {tr_in_end, tr_k_first, tr_k_last, tr_out_first, tr_out_last}) cudaEventDestroy(ev);
  }
  for (int g = 0; g < ngroups; ++g) {
    cudaEventDestroy(ev_in[g]);
    cudaEventDestroy(ev_done[g]);
  }
  cudaEventDestroy(ev_zero);
  if (result) return result;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
    return fail(MLX_ERR_CUDA, std::string("pipeline failed: ") +
                                  cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
  return MLX_OK;
}

int mlx_pv_process_host(mlx_ctx* c, const mlx_pv_params* p, const float* const* wav, const int64_t* n, int ntracks,
                        float* const* out_wav, int32_t* const* out_peak, float* const* out_f0) {
  return pv_process_host_impl(c, p, reinterpret_cast<const void* const*>(wav), MLX_FMT_F32, n, ntracks,
                              reinterpret_cast<void* const*>(out_wav), MLX_FMT_F32, out_peak, out_f0);
}

int mlx_pv_process_host_fmt(mlx_ctx* c, const mlx_pv_params* p, const void* const* wav, int in_format,
                            const int64_t* n, int ntracks, void* const* out_wav, int out_format,
                            int32_t* const* out_peak, float* const* out_f0) {
  return pv_process_host_impl(c, p, wav, in_format, n, ntracks, out_wav, out_format, out_peak, out_f0);
}

// ------------------------------------------------------------------------------------------------ grains
int mlx_grain_render(mlx_ctx* c, int track, const int32_t* g_start, const int32_t* g_len, const float* g_rate,
                     const int64_t* out_off, const float* g_next, int ngrains, int tail_zeros, float* out,
                     int16_t* out_i16) {
  if (!c || ngrains < 0 || tail_zeros < 0) return fail(MLX_ERR_INVALID, "bad argument");
  if (track < 0 || track >= (int)c->tracks.size()) return fail(MLX_ERR_STATE, "track not uploaded");
  if (ngrains > 0 && (!g_start || !g_len || !g_rate || !out_off || !g_next))
    return fail(MLX_ERR_INVALID, "null schedule");
  CK(cudaSetDevice(c->device));
  const int64_t n = c->tracks[track].n;
  int64_t zero_off = 0;
  const int64_t rendered = ngrains > 0 ? out_off[ngrains] : 0;
  for (int g = 0; g < ngrains; ++g) {
    if (g_start[g] < 0 || g_len[g] <= 0 || (int64_t)g_start[g] + g_len[g] > n || out_off[g + 1] < out_off[g])
      return fail(MLX_ERR_INVALID, "schedule row outside the track");
    // every sample a row renders must index inside its grain (reference stops at idx >= size)
    const int64_t cnt = out_off[g + 1] - out_off[g];
    if (cnt > 0 && (int64_t)truncf((float)(cnt - 1) * g_rate[g]) >= g_len[g])
      return fail(MLX_ERR_INVALID, "schedule row renders past its grain");
  }
  const int64_t total = rendered + tail_zeros;
  if (total == 0) return MLX_OK;
  const size_t ng1 = (size_t)std::max(ngrains, 1);
  CK(c->g_i32a.reserve(sizeof(int32_t) * ng1));
  CK(c->g_i32b.reserve(sizeof(int32_t) * ng1));
  CK(c->g_f32a.reserve(sizeof(float) * ng1));
  CK(c->g_f32b.reserve(sizeof(float) * ng1));
  CK(c->g_i64.reserve(sizeof(int64_t) * (ng1 + 1)));
  if (out) CK(c->g_out.reserve(sizeof(float) * total));
  if (out_i16) CK(c->g_out16.reserve(sizeof(int16_t) * total));
  if (ngrains > 0) {
    CK(cudaMemcpyAsync(c->g_i32a.p, g_start, sizeof(int32_t) * ngrains, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->g_i32b.p, g_len, sizeof(int32_t) * ngrains, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->g_f32a.p, g_rate, sizeof(float) * ngrains, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->g_f32b.p, g_next, sizeof(float) * ngrains, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->g_i64.p, out_off, sizeof(int64_t) * (ngrains + 1), cudaMemcpyHostToDevice, c->stream));
  } else {
    CK(cudaMemcpyAsync(c->g_i64.p, &zero_off, sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  }
  GrainArgs a{};
  a.x = c->track_ptr(track);
  a.g_start = static_cast<const int*>(c->g_i32a.p);
  a.g_len = static_cast<const int*>(c->g_i32b.p);
  a.g_rate = static_cast<const float*>(c->g_f32a.p);
  a.g_next = static_cast<const float*>(c->g_f32b.p);
  a.out_off = static_cast<const long long*>(c->g_i64.p);
  a.ngrains = ngrains;
  a.total = total;
  a.out = out ? static_cast<float*>(c->g_out.p) : nullptr;
  a.out_i16 = out_i16 ? static_cast<short*>(c->g_out16.p) : nullptr;
  c->mark(4);
  CK(launch_grain(a, c->stream));
  c->mark(-1);
  c->launches += 1;
  if (out) CK(cudaMemcpyAsync(out, c->g_out.p, sizeof(float) * total, cudaMemcpyDeviceToHost, c->stream));
  if (out_i16)
    CK(cudaMemcpyAsync(out_i16, c->g_out16.p, sizeof(int16_t) * total, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

// Grain segmentation of every uploaded track (K8a crossings + K8b chain, grain_kernels.cu).
// Row t of the outputs has `cap` entries; counts[t] may exceed cap (then only cap rows were written).
static int grain_segment_common(mlx_ctx* c, int32_t* const* g_start_dev, int32_t* const* g_len_dev, int cap,
                                int32_t* counts_dev) {
  const int nt = (int)c->tracks.size();
  // bit arrays: two per track, ceil(n / 32) + 1 words each, 16-byte aligned
  std::vector<size_t> woff(nt + 1, 0);
  long long max_words = 0;
  for (int t = 0; t < nt; ++t) {
    const long long nw = (c->tracks[t].n + 31) / 32 + 1;
    max_words = std::max(max_words, nw);
    woff[t + 1] = woff[t] + 2 * (size_t)((nw + 3) & ~3LL);
  }
  CK(c->seg_bits.reserve(sizeof(uint32_t) * std::max<size_t>(woff[nt], 4)));
  CK(c->seg_desc.reserve(sizeof(GrainSegTrack) * nt));
  mlx_ctx::Slot* slot = nullptr;
  int rc = acquire_slot(c, sizeof(GrainSegTrack) * nt, &slot);
  if (rc) return rc;
  GrainSegTrack* d = static_cast<GrainSegTrack*>(slot->p);
  uint32_t* bits = static_cast<uint32_t*>(c->seg_bits.p);
  for (int t = 0; t < nt; ++t) {
    const long long nw = (c->tracks[t].n + 31) / 32 + 1;
    d[t].x = c->track_ptr(t);
    d[t].n = c->tracks[t].n;
    d[t].nwords = nw;
    d[t].z7 = bits + woff[t];
    d[t].z3 = bits + woff[t] + (size_t)((nw + 3) & ~3LL);
    d[t].g_start = g_start_dev[t];
    d[t].g_len = g_len_dev[t];
    d[t].count = counts_dev + t;
  }
  CK(cudaMemcpyAsync(c->seg_desc.p, d, sizeof(GrainSegTrack) * nt, cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(slot->done, c->stream));
  c->mark(4);
  CK(launch_grain_segment(static_cast<const GrainSegTrack*>(c->seg_desc.p), nt, max_words, cap, c->stream));
  c->mark(-1);
  c->launches += 2;
  return MLX_OK;
}

int mlx_grain_segment_dev(mlx_ctx* c, int32_t* const* g_start_dev, int32_t* const* g_len_dev, int cap,
                          int32_t* counts_dev) {
  if (!c || !g_start_dev || !g_len_dev || !counts_dev || cap < 0) return fail(MLX_ERR_INVALID, "bad argument");
  if (c->tracks.empty()) return fail(MLX_ERR_STATE, "no tracks uploaded");
  for (size_t t = 0; t < c->tracks.size(); ++t)
    if (cap > 0 && (!g_start_dev[t] || !g_len_dev[t])) return fail(MLX_ERR_INVALID, "null output row");
  CK(cudaSetDevice(c->device));
  return grain_segment_common(c, g_start_dev, g_len_dev, cap, counts_dev);
}

int mlx_grain_segment(mlx_ctx* c, int32_t* const* g_start, int32_t* const* g_len, int cap, int32_t* counts) {
  if (!c || !g_start || !g_len || !counts || cap < 0) return fail(MLX_ERR_INVALID, "bad argument");
  if (c->tracks.empty()) return fail(MLX_ERR_STATE, "no tracks uploaded");
  const int nt = (int)c->tracks.size();
  for (int t = 0; t < nt; ++t)
    if (cap > 0 && (!g_start[t] || !g_len[t])) return fail(MLX_ERR_INVALID, "null output row");
  CK(cudaSetDevice(c->device));
  const size_t row = (size_t)std::max(cap, 1);
  CK(c->seg_rows.reserve(sizeof(int32_t) * 2 * row * nt));
  CK(c->seg_count.reserve(sizeof(int32_t) * nt));
  int32_t* rows = static_cast<int32_t*>(c->seg_rows.p);
  std::vector<int32_t*> ps(nt), pl(nt);
  for (int t = 0; t < nt; ++t) {
    ps[t] = rows + (size_t)(2 * t) * row;
    pl[t] = rows + (size_t)(2 * t + 1) * row;
  }
  int rc = grain_segment_common(c, ps.data(), pl.data(), cap, static_cast<int32_t*>(c->seg_count.p));
  if (rc) return rc;
  CK(cudaMemcpyAsync(counts, c->seg_count.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int t = 0; t < nt; ++t) {
    const int m = std::min(counts[t], cap);
    if (m <= 0) continue;
    CK(cudaMemcpyAsync(g_start[t], ps[t], sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(g_len[t], pl[t], sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

// ------------------------------------------------------------------------------------------------ picks
int mlx_picks_levels(int64_t n) {
  int lvl = 0;
  while (n > ((int64_t)1 << (lvl + 1))) ++lvl;  // reference app.cpp:352, :365
  return lvl;
}

int64_t mlx_picks_layout(int64_t n, int64_t* level_off) {
  const int L = mlx_picks_levels(n);
  int64_t off = 0;
  for (int l = 0; l < L; ++l) {
    if (level_off) level_off[l] = off;
    off += n >> (l + 1);  // reference app.cpp:356, :369
  }
  if (level_off) level_off[L] = off;
  return off;
}

static int picks_fill_args(mlx_ctx* c, int track, float* pairs_dev, PicksArgs* a) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  if (track < 0 || track >= (int)c->tracks.size()) return fail(MLX_ERR_STATE, "track not uploaded");
  a->x = c->track_ptr(track);
  a->n = c->tracks[track].n;
  a->levels = mlx_picks_levels(a->n);
  int64_t off[33];
  mlx_picks_layout(a->n, off);
  for (int l = 0; l <= a->levels; ++l) a->level_off[l] = off[l];
  a->pairs = reinterpret_cast<float2*>(pairs_dev);
  return MLX_OK;
}

// pyramids of tracks [first, first + count) in one launch; pairs_dev[i] belongs to track first + i
static int picks_build_range(mlx_ctx* c, int first, int count, float* const* pairs_dev) {
  CK(cudaSetDevice(c->device));
  CK(c->picks_desc.reserve(sizeof(PicksArgs) * count));
  mlx_ctx::Slot* slot = nullptr;
  int rc = acquire_slot(c, sizeof(PicksArgs) * count, &slot);
  if (rc) return rc;
  PicksArgs* d = static_cast<PicksArgs*>(slot->p);
  long long max_n = 0;
  int max_levels = 0;
  for (int i = 0; i < count; ++i) {
    rc = picks_fill_args(c, first + i, pairs_dev[i], &d[i]);
    if (rc) return rc;
    if (d[i].levels > 0 && !pairs_dev[i]) return fail(MLX_ERR_INVALID, "null argument");
    max_n = std::max(max_n, d[i].n);
    max_levels = std::max(max_levels, d[i].levels);
  }
  CK(cudaMemcpyAsync(c->picks_desc.p, d, sizeof(PicksArgs) * count, cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(slot->done, c->stream));
  c->mark(4);
  CK(launch_picks_build(static_cast<const PicksArgs*>(c->picks_desc.p), count, max_n, max_levels, c->stream));
  c->mark(-1);
  if (max_levels > 0) c->launches += max_levels > 12 ? 2 : 1;
  return MLX_OK;
}

int mlx_picks_build_dev(mlx_ctx* c, int track, float* pairs_dev) {
  if (!c) return fail(MLX_ERR_INVALID, "ctx is null");
  if (track < 0 || track >= (int)c->tracks.size()) return fail(MLX_ERR_STATE, "track not uploaded");
  return picks_build_range(c, track, 1, &pairs_dev);
}

int mlx_picks_build_all_dev(mlx_ctx* c, float* const* pairs_dev) {
  if (!c || !pairs_dev) return fail(MLX_ERR_INVALID, "null argument");
  if (c->tracks.empty()) return fail(MLX_ERR_STATE, "no tracks uploaded");
  return picks_build_range(c, 0, (int)c->tracks.size(), pairs_dev);
}

// builds (once per upload and track) the pyramid the host entry points work on
static int picks_ensure(mlx_ctx* c, int track, PicksArgs* a) {
  int rc = picks_fill_args(c, track, nullptr, a);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  const size_t total = (size_t)a->level_off[a->levels];
  if (c->picks_track != track) {
    CK(c->picks.reserve(sizeof(float) * 2 * std::max<size_t>(total, 1)));
    rc = mlx_picks_build_dev(c, track, static_cast<float*>(c->picks.p));
    if (rc) return rc;
    c->picks_track = track;
  }
  a->pairs = static_cast<float2*>(c->picks.p);
  return MLX_OK;
}

int mlx_picks_build(mlx_ctx* c, int track, float* pairs) {
  PicksArgs a{};
  int rc = picks_ensure(c, track, &a);
  if (rc) return rc;
  const size_t total = (size_t)a.level_off[a.levels];
  if (total == 0) return MLX_OK;
  if (!pairs) return fail(MLX_ERR_INVALID, "null argument");
  CK(cudaMemcpyAsync(pairs, c->picks.p, sizeof(float) * 2 * total, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

int mlx_minmax_ranges(mlx_ctx* c, int track, const int32_t* start_end, int count, float* out) {
  if (count < 0 || (count > 0 && (!start_end || !out))) return fail(MLX_ERR_INVALID, "bad argument");
  PicksArgs a{};
  int rc = picks_ensure(c, track, &a);
  if (rc) return rc;
  if (count == 0) return MLX_OK;
  CK(c->picks_ranges.reserve(sizeof(int32_t) * 2 * (size_t)count));
  CK(c->picks_out.reserve(sizeof(float) * 2 * (size_t)count));
  CK(cudaMemcpyAsync(c->picks_ranges.p, start_end, sizeof(int32_t) * 2 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
  c->mark(4);
  CK(launch_minmax_ranges(a, static_cast<const int*>(c->picks_ranges.p), count, static_cast<float*>(c->picks_out.p),
                          c->stream));
  c->mark(-1);
  c->launches += 1;
  CK(cudaMemcpyAsync(out, c->picks_out.p, sizeof(float) * 2 * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

}  // extern "C"
