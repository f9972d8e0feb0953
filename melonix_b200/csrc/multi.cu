// melonix_b200/csrc/multi.cu -- multi-GPU layer behind the C ABI: one long file phase-vocoded by
// contiguous time range, one rank (process or thread) per GPU (BASELINE configs[3]).
//
// Frame convention as everywhere (reference spec.cpp:47, spec-cache.cpp:63-65): frame f covers samples
// [(f+1) hop - fftN, (f+1) hop).  Rank r owns frames [F r / G, F (r+1) / G) and the output hops of those
// frames.  Per run and rank:
//   1. seam exchange -- ONE NCCL group of send/recv with the two neighbours: the samples before the
//      owned range that its first frames (and the halo frame that seeds the phase difference) reach back
//      into, and the samples after it that its last three overlap-add frames reach into.  Received
//      straight into the zero-padded track buffer the kernels read (no staging copy); the copy of the
//      rank's own samples into that buffer runs on the compute stream while the exchange is in flight.
//   2. K_A once (mlx_pv_analyze_dev): intermediates stay staged, per-bin uint32 phase totals result.
//   3. all-gather of the totals (fftN/2+32 words per track and rank) + a prefix over the lower ranks:
//      integer addition is associative, so the carried-in phase is exact and the sharded output is
//      bit-identical to the unsharded run.
//   4. scan + K_S on the staged analysis (mlx_pv_synth_dev) writing only the owned hops.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): libmelonix_b200.so keeps no link-time dependency
// on it, single-GPU users never load it, and inside a PyTorch process the already-loaded NCCL is reused.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "capi_internal.h"

using namespace mlx;

namespace {

// the few NCCL entry points used, with their ABI-stable signatures (nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint32 = 3, ncclFloat32 = 7 };  // ncclDataType_t values (nccl.h: ncclUint32 = 3, ncclFloat32 = 7)

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string err;
};

Nccl* nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("MLX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) break;
      n.err = dlerror();
    }
    if (!n.h) return;
    auto sym = [&](const char* s) {
      void* p = dlsym(n.h, s);
      if (!p) n.err = std::string("missing symbol ") + s;
      return p;
    };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.Send = reinterpret_cast<decltype(n.Send)>(sym("ncclSend"));
    n.Recv = reinterpret_cast<decltype(n.Recv)>(sym("ncclRecv"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.Send || !n.Recv || !n.AllGather || !n.GroupStart ||
        !n.GroupEnd || !n.GetErrorString) {
      dlclose(n.h);
      n.h = nullptr;
    }
  });
  return n.h ? &n : nullptr;
}

#define NK(expr)                                                                                  \
  do {                                                                                            \
    int r_ = (expr);                                                                              \
    if (r_ != ncclSuccess) return fail(MLX_ERR_CUDA, std::string(#expr) + ": " + N->GetErrorString(r_)); \
  } while (0)

// carry[t][j] = sum over ranks r < rank of gathered[r][t][j]  (mod 2^32)
__global__ void phase_prefix_kernel(const uint32_t* __restrict__ gathered, int rank, size_t per_rank, uint32_t* carry) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_rank) return;
  uint32_t acc = 0u;
  for (int r = 0; r < rank; ++r) acc += gathered[(size_t)r * per_rank + i];
  carry[i] = acc;
}

}  // namespace

struct mlx_comm {
  mlx_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_seam = nullptr;
  DevBuf totals, gathered, carry;
  DevBuf h_in, h_out, h_peak, h_f0;  // device staging of the host-pointer entry point
};

extern "C" {

int mlx_shard_frames(int64_t n, int fftN, int hop, int world, int rank, mlx_time_shard* out) {
  if (!out || n < 0 || hop <= 0 || fftN != 4 * hop || world <= 0 || rank < 0 || rank >= world)
    return fail(MLX_ERR_INVALID, "mlx_shard_frames: bad argument");
  const int64_t F = (n + hop - 1) / hop;
  const int64_t fb = F * rank / world, fe = F * (rank + 1) / world;
  // frame f covers samples [(f-3) hop, (f+1) hop).  Needed frames: fb-1 (the halo frame whose spectrum
  // seeds the phase difference) .. fe+2 (their tails overlap-add into the owned hops).
  const int64_t off = std::max<int64_t>(fb - 4, 0);
  out->frame_begin = fb;
  out->frame_end = fe;
  out->frame_offset = off;
  out->need_lo = off * hop;
  out->need_hi = std::min<int64_t>(n, (fe + 3) * hop);
  out->own_lo = std::min<int64_t>(n, fb * hop);
  out->own_hi = std::min<int64_t>(n, fe * hop);
  return MLX_OK;
}

int mlx_comm_unique_id(void* id128) {
  if (!id128) return fail(MLX_ERR_INVALID, "id128 is null");
  Nccl* N = nccl();
  if (!N) return fail(MLX_ERR_UNSUPPORTED, "NCCL is not loadable (dlopen libnccl.so.2; override with MLX_NCCL_LIB)");
  ncclUniqueId id;
  NK(N->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return MLX_OK;
}

int mlx_comm_create(mlx_comm** out, mlx_ctx* ctx, const void* id128, int world, int rank) {
  if (!out || !ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(MLX_ERR_INVALID, "bad argument");
  *out = nullptr;
  Nccl* N = nccl();
  if (!N) return fail(MLX_ERR_UNSUPPORTED, "NCCL is not loadable (dlopen libnccl.so.2; override with MLX_NCCL_LIB)");
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  mlx_comm* m = new mlx_comm();
  m->ctx = ctx;
  m->world = world;
  m->rank = rank;
  int r = N->CommInitRank(&m->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete m;
    return fail(MLX_ERR_CUDA, std::string("ncclCommInitRank: ") + N->GetErrorString(r));
  }
  if (cudaStreamCreateWithFlags(&m->s_comm, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&m->ev_seam, cudaEventDisableTiming) != cudaSuccess) {
    N->CommDestroy(m->comm);
    delete m;
    return fail(MLX_ERR_CUDA, "stream / event creation failed");
  }
  *out = m;
  return MLX_OK;
}

void mlx_comm_destroy(mlx_comm* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->s_comm);
  if (Nccl* N = nccl()) N->CommDestroy(m->comm);
  cudaStreamDestroy(m->s_comm);
  cudaEventDestroy(m->ev_ready);
  cudaEventDestroy(m->ev_seam);
  m->totals.release();
  m->gathered.release();
  m->carry.release();
  for (DevBuf* b : {&m->h_in, &m->h_out, &m->h_peak, &m->h_f0}) b->release();
  delete m;
}

int mlx_comm_info(const mlx_comm* m, int* world, int* rank, int* nccl_version) {
  if (!m) return fail(MLX_ERR_INVALID, "comm is null");
  if (world) *world = m->world;
  if (rank) *rank = m->rank;
  if (nccl_version) {
    *nccl_version = 0;
    Nccl* N = nccl();
    if (N && N->GetVersion) N->GetVersion(nccl_version);
  }
  return MLX_OK;
}

int mlx_pv_run_sharded_dev(mlx_ctx* c, mlx_comm* m, const mlx_pv_params* p, const float* const* own_dev, int ntracks,
                           int64_t n_total, float* const* out_own_dev, int32_t* const* peak_own_dev,
                           float* const* f0_own_dev) {
  if (!c || !m || !p || !own_dev || ntracks <= 0 || m->ctx != c) return fail(MLX_ERR_INVALID, "bad argument");
  if (p->frame_begin > 0 || p->frame_end >= 0 || p->phase_in_dev || p->rate_per_frame_dev)
    return fail(MLX_ERR_UNSUPPORTED, "mlx_pv_run_sharded_dev shards the whole file itself (constant rate, no frame range)");
  Nccl* N = nccl();
  if (!N) return fail(MLX_ERR_UNSUPPORTED, "NCCL is not loadable");
  CK(cudaSetDevice(c->device));
  const int world = m->world, rank = m->rank, fftN = p->fftN, hop = p->hop;
  mlx_time_shard me, left, right;
  int rc = mlx_shard_frames(n_total, fftN, hop, world, rank, &me);
  if (rc) return rc;
  if (rank > 0 && (rc = mlx_shard_frames(n_total, fftN, hop, world, rank - 1, &left))) return rc;
  if (rank + 1 < world && (rc = mlx_shard_frames(n_total, fftN, hop, world, rank + 1, &right))) return rc;
  const int64_t own = me.own_hi - me.own_lo, lh = me.own_lo - me.need_lo, rh = me.need_hi - me.own_hi;
  const int64_t to_right = rank + 1 < world ? right.own_lo - right.need_lo : 0;  // my tail -> its left halo
  const int64_t to_left = rank > 0 ? left.need_hi - left.own_hi : 0;             // my head -> its right halo
  if (to_right > own || to_left > own)
    return fail(MLX_ERR_UNSUPPORTED, "time shards are shorter than the seam overlap: use fewer ranks for this file");

  // 1. lay the local windows out (zero padding around them) and start the seam exchange
  std::vector<int64_t> nloc(ntracks, me.need_hi - me.need_lo);
  rc = layout_tracks(c, nloc.data(), ntracks);  // memset of the buffer is queued on c->stream
  if (rc) return rc;
  CK(cudaEventRecord(m->ev_ready, c->stream));
  CK(cudaStreamWaitEvent(m->s_comm, m->ev_ready, 0));  // halos land after the zero fill
  NK(N->GroupStart());
  for (int t = 0; t < ntracks; ++t) {
    float* base = static_cast<float*>(c->track_buf.p) + c->tracks[t].offset;
    if (rank > 0 && lh > 0) NK(N->Recv(base, (size_t)lh, ncclFloat32, rank - 1, m->comm, m->s_comm));
    if (rank + 1 < world && rh > 0) NK(N->Recv(base + lh + own, (size_t)rh, ncclFloat32, rank + 1, m->comm, m->s_comm));
    if (to_right > 0) NK(N->Send(own_dev[t] + (own - to_right), (size_t)to_right, ncclFloat32, rank + 1, m->comm, m->s_comm));
    if (to_left > 0) NK(N->Send(own_dev[t], (size_t)to_left, ncclFloat32, rank - 1, m->comm, m->s_comm));
  }
  NK(N->GroupEnd());
  CK(cudaEventRecord(m->ev_seam, m->s_comm));
  // ... while the rank's own samples (the bulk) are copied into place on the compute stream
  for (int t = 0; t < ntracks; ++t)
    if (own > 0)
      CK(cudaMemcpyAsync(static_cast<float*>(c->track_buf.p) + c->tracks[t].offset + lh, own_dev[t], sizeof(float) * own,
                         cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaStreamWaitEvent(c->stream, m->ev_seam, 0));

  // 2. analysis of the owned frames (local indices), totals per (track, bin)
  const int NBP = pv_nbp(fftN), NB = fftN / 2 + 1;
  const size_t per_rank = (size_t)ntracks * NBP;
  CK(m->totals.reserve(sizeof(uint32_t) * per_rank));
  CK(m->gathered.reserve(sizeof(uint32_t) * per_rank * world));
  CK(m->carry.reserve(sizeof(uint32_t) * per_rank));
  CK(cudaMemsetAsync(m->totals.p, 0, sizeof(uint32_t) * per_rank, c->stream));
  mlx_pv_params q = *p;
  q.frame_begin = me.frame_begin - me.frame_offset;
  q.frame_end = me.frame_end - me.frame_offset;
  q.wave_mib = -1;
  const int64_t lfb = q.frame_begin;
  std::vector<uint32_t*> tot(ntracks);
  std::vector<int32_t*> pk(ntracks, nullptr);
  std::vector<float*> f0(ntracks, nullptr), yw(ntracks, nullptr);
  for (int t = 0; t < ntracks; ++t) {
    tot[t] = static_cast<uint32_t*>(m->totals.p) + (size_t)t * NBP;
    // outputs are indexed by LOCAL frame / sample; only owned frames / hops are written, so the caller's
    // owned-range arrays are addressed through a shifted base pointer
    if (peak_own_dev && peak_own_dev[t]) pk[t] = peak_own_dev[t] - lfb;
    if (f0_own_dev && f0_own_dev[t]) f0[t] = f0_own_dev[t] - lfb;
    if (out_own_dev && out_own_dev[t]) yw[t] = out_own_dev[t] - lh;
  }
  rc = mlx_pv_analyze_dev(c, &q, tot.data(), peak_own_dev ? pk.data() : nullptr, f0_own_dev ? f0.data() : nullptr);
  if (rc) return rc;

  // 3. phase carry: all-gather of the totals, exclusive prefix over the lower ranks
  if (world > 1) {
    NK(N->AllGather(m->totals.p, m->gathered.p, per_rank, ncclUint32, m->comm, c->stream));
    phase_prefix_kernel<<<(unsigned)((per_rank + 255) / 256), 256, 0, c->stream>>>(
        static_cast<const uint32_t*>(m->gathered.p), rank, per_rank, static_cast<uint32_t*>(m->carry.p));
    CK(cudaGetLastError());
    c->launches += 1;
  } else {
    CK(cudaMemsetAsync(m->carry.p, 0, sizeof(uint32_t) * per_rank, c->stream));
  }
  (void)NB;

  // 4. synthesis of the owned hops from the staged analysis
  if (!out_own_dev) return MLX_OK;
  std::vector<const uint32_t*> cin(ntracks);
  for (int t = 0; t < ntracks; ++t) cin[t] = static_cast<const uint32_t*>(m->carry.p) + (size_t)t * NBP;
  q.phase_in_dev = cin.data();
  return mlx_pv_synth_dev(c, &q, yw.data());
}

// Host-pointer form for a C++ host (the reference's language): uploads the owned samples, runs the
// sharded pipeline, downloads the owned results.  Blocking.
int mlx_pv_run_sharded(mlx_ctx* c, mlx_comm* m, const mlx_pv_params* p, const float* const* own, int ntracks,
                       int64_t n_total, float* const* out_own, int32_t* const* peak_own, float* const* f0_own) {
  if (!c || !m || !p || !own || ntracks <= 0) return fail(MLX_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  mlx_time_shard me;
  int rc = mlx_shard_frames(n_total, p->fftN, p->hop, m->world, m->rank, &me);
  if (rc) return rc;
  const size_t own_n = (size_t)(me.own_hi - me.own_lo), nf = (size_t)(me.frame_end - me.frame_begin);
  const size_t own_pad = (own_n + 31) & ~size_t(31), nf_pad = (nf + 31) & ~size_t(31);
  CK(m->h_in.reserve(sizeof(float) * std::max<size_t>(own_pad * ntracks, 1)));
  CK(m->h_out.reserve(sizeof(float) * std::max<size_t>(own_pad * ntracks, 1)));
  CK(m->h_peak.reserve(sizeof(int32_t) * std::max<size_t>(nf_pad * ntracks, 1)));
  CK(m->h_f0.reserve(sizeof(float) * std::max<size_t>(nf_pad * ntracks, 1)));
  std::vector<const float*> din(ntracks);
  std::vector<float*> dout(ntracks), df0(ntracks);
  std::vector<int32_t*> dpk(ntracks);
  for (int t = 0; t < ntracks; ++t) {
    din[t] = static_cast<float*>(m->h_in.p) + own_pad * t;
    dout[t] = static_cast<float*>(m->h_out.p) + own_pad * t;
    dpk[t] = static_cast<int32_t*>(m->h_peak.p) + nf_pad * t;
    df0[t] = static_cast<float*>(m->h_f0.p) + nf_pad * t;
    if (own_n) CK(cudaMemcpyAsync(const_cast<float*>(din[t]), own[t], sizeof(float) * own_n, cudaMemcpyHostToDevice, c->stream));
  }
  rc = mlx_pv_run_sharded_dev(c, m, p, din.data(), ntracks, n_total, dout.data(), dpk.data(), df0.data());
  if (rc) return rc;
  for (int t = 0; t < ntracks; ++t) {
    if (out_own && out_own[t] && own_n)
      CK(cudaMemcpyAsync(out_own[t], dout[t], sizeof(float) * own_n, cudaMemcpyDeviceToHost, c->stream));
    if (peak_own && peak_own[t] && nf)
      CK(cudaMemcpyAsync(peak_own[t], dpk[t], sizeof(int32_t) * nf, cudaMemcpyDeviceToHost, c->stream));
    if (f0_own && f0_own[t] && nf)
      CK(cudaMemcpyAsync(f0_own[t], df0[t], sizeof(float) * nf, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return MLX_OK;
}

}  // extern "C"
