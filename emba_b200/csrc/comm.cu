// Combination of the per-GPU partial systems over NVLink: NCCL all-reduce of the
// num_ev_map histogram, the cost scalars, A11/b1, A22/b2 and the A12 strips on the handle's stream. NCCL is
// loaded lazily (dlopen) so that the library has no hard dependency on it for single-GPU use.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <vector>

#include <cstring>

#include <cuda.h>

#include "emba_internal.cuh"

namespace emba {

typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_get_uid)(nccl_uid_t*);
typedef int (*fn_init_rank)(void**, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_sendrecv send = nullptr;
  fn_sendrecv recv = nullptr;
  fn_group group_start = nullptr;
  fn_group group_end = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.get_uid = (fn_get_uid)dlsym(api.lib, "ncclGetUniqueId");
  api.init_rank = (fn_init_rank)dlsym(api.lib, "ncclCommInitRank");
  api.allreduce = (fn_allreduce)dlsym(api.lib, "ncclAllReduce");
  api.allgather = (fn_allgather)dlsym(api.lib, "ncclAllGather");
  api.send = (fn_sendrecv)dlsym(api.lib, "ncclSend");
  api.recv = (fn_sendrecv)dlsym(api.lib, "ncclRecv");
  api.group_start = (fn_group)dlsym(api.lib, "ncclGroupStart");
  api.group_end = (fn_group)dlsym(api.lib, "ncclGroupEnd");
  api.destroy = (fn_destroy)dlsym(api.lib, "ncclCommDestroy");
  api.errstr = (fn_errstr)dlsym(api.lib, "ncclGetErrorString");
  if (!api.get_uid || !api.init_rank || !api.allreduce || !api.destroy || !api.allgather || !api.send || !api.recv ||
      !api.group_start || !api.group_end) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

// dtype: 0 = int32 sum, 1 = fp64 sum, 2 = int32 min, 3 = int32 max
int comm_allreduce(Handle* h, void* buf, int64_t count, int dtype) {
  if (h->world <= 1 || count <= 0) return EMBA_OK;
  NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) { h->err = "multi-GPU shard without a communicator: call emba_comm_init"; return EMBA_E_NCCL; }
  // ncclDataType_t: ncclInt32 = 2, ncclFloat64 = 8; ncclRedOp_t: sum 0, prod 1, max 2, min 3
  const int ty = (dtype == 1) ? 8 : 2;
  const int op = (dtype == 2) ? 3 : (dtype == 3) ? 2 : 0;
  const int r = api->allreduce(buf, buf, (size_t)count, ty, op, h->nccl_comm, h->stream);
  if (r != 0) { h->err = std::string("ncclAllReduce: ") + (api->errstr ? api->errstr(r) : "error"); return EMBA_E_NCCL; }
  h->launches++;
  return EMBA_OK;
}

int comm_allgather_i32(Handle* h, const int32_t* send, int32_t* recv, int64_t count) {
  if (h->world <= 1) return EMBA_OK;
  NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) { h->err = "multi-GPU shard without a communicator: call emba_comm_init before emba_set_events"; return EMBA_E_NCCL; }
  if (api->allgather(send, recv, (size_t)count, 2, h->nccl_comm, h->stream) != 0) { h->err = "ncclAllGather failed"; return EMBA_E_NCCL; }
  h->launches++;
  return EMBA_OK;
}

// the evaluation's reductions as ONE NCCL group: int32 histogram over the panorama (rank-local counts in, global
// counts out -- the local ones size the rank's row segments afterwards), fp64 cost and count, and the range flag
// (max), so that every rank takes the same decision on it
int comm_allreduce_eval(Handle* h, const int32_t* hist_loc, int32_t* hist, int64_t P, double* scal2, int32_t* flags) {
  if (h->world <= 1) return EMBA_OK;
  NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) { h->err = "multi-GPU shard without a communicator: call emba_comm_init"; return EMBA_E_NCCL; }
  if (api->group_start() != 0) { h->err = "ncclGroupStart failed"; return EMBA_E_NCCL; }
  int rc = api->allreduce(hist_loc, hist, (size_t)P, 2, 0, h->nccl_comm, h->stream);
  rc |= api->allreduce(scal2, scal2, 2, 8, 0, h->nccl_comm, h->stream);
  rc |= api->allreduce(flags, flags, 1, 2, 2, h->nccl_comm, h->stream);
  if (api->group_end() != 0 || rc != 0) { h->err = "ncclAllReduce (histogram, cost) failed"; return EMBA_E_NCCL; }
  h->launches++;
  return EMBA_OK;
}

// ---------------------------------------------------------------------------------------------------
// A12 exchange for the time-sharded path. Rank r holds, for every active pixel a, the sub-strip of the control
// poses its own time slice touches (local window [lo_r(a), hi_r(a)]). Rank q owns the contiguous pixel range
// [Np q / W, Np (q+1) / W). Because local strips are stored in pixel order, the sub-strips destined to q are ONE
// contiguous chunk of the local strip buffer: the exchange is a plain all-to-all of contiguous chunks
// (ncclSend/ncclRecv over NVLink), 1/W of the A12 volume per rank, no packing. The owner then merges the W
// sub-strips of each of its pixels into the strip over the merged window, in rank order (deterministic).
// ---------------------------------------------------------------------------------------------------
__global__ void k_pack_win(int64_t Np, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                           int32_t* __restrict__ out) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  out[2 * a] = lo[a];
  out[2 * a + 1] = hi[a];
}

// per source rank s: lengths of the sub-strips of my pixels; merged windows and lengths
__global__ void k_own_len(int W, int64_t Np, int64_t a0, int64_t n_own, const int32_t* __restrict__ win_all,
                          int64_t* __restrict__ own_len, int32_t* __restrict__ gwinlo, int32_t* __restrict__ gwinhi,
                          int64_t* __restrict__ glen) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a > Np) return;
  if (a == Np) { glen[Np] = 0; return; }
  const bool mine = a >= a0 && a < a0 + n_own;
  int glo = INT_MAX, ghi = -1;
  if (mine) {
    for (int s = 0; s < W; s++) {
      const int lo = win_all[((size_t)s * Np + a) * 2], hi = win_all[((size_t)s * Np + a) * 2 + 1];
      const int64_t len = hi >= lo ? (int64_t)(hi - lo + 1) : 0;
      own_len[(size_t)s * (n_own + 1) + (a - a0)] = len;
      if (len > 0) { glo = min(glo, lo); ghi = max(ghi, hi); }
    }
  }
  gwinlo[a] = glo;
  gwinhi[a] = ghi;
  glen[a] = ghi >= glo ? (int64_t)(ghi - glo + 1) : 0;
}

__global__ void k_zero_tail(int W, int64_t n_own, int64_t* __restrict__ own_len) {
  const int s = threadIdx.x;
  if (s < W) own_len[(size_t)s * (n_own + 1) + n_own] = 0;
}

// one warp per owned pixel: strip over the merged window = sum over source ranks (fixed order) of their sub-strips.
// The per-source windows and receive offsets are loaded once per pixel (not per element), elements move as
// double2; MAXW > 0 keeps them in registers for worlds up to MAXW ranks, MAXW == 0 is the generic loop.
template <int MAXW>
__global__ void k_merge_strips(int W, int64_t Np, int64_t a0, int64_t n_own, const int32_t* __restrict__ win_all,
                               const int64_t* __restrict__ own_off,
                               const double* __restrict__ recv, const int32_t* __restrict__ gwinlo,
                               const int64_t* __restrict__ gstripoff, double* __restrict__ gstrip, int group,
                               unsigned long long* __restrict__ gmask) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_own) return;
  const int64_t a = a0 + i;
  const int glo = gwinlo[a];
  const int64_t goff = gstripoff[a];
  const int64_t glen = gstripoff[a + 1] - goff;
  constexpr int NW = MAXW > 0 ? MAXW : 1;
  int lo[NW], hi[NW];
  int64_t base[NW];  // double2 index of the source's sub-strip in the receive buffer
  if (MAXW > 0) {
#pragma unroll
    for (int s = 0; s < NW; s++) {
      lo[s] = INT_MAX; hi[s] = -1; base[s] = 0;
      if (s < W) {
        lo[s] = win_all[((size_t)s * Np + a) * 2];
        hi[s] = win_all[((size_t)s * Np + a) * 2 + 1];
        base[s] = own_off[(size_t)s * (n_own + 1) + i] * 3;  // flattened scan: receive-chunk base included
      }
    }
  }
  const double2* recv2 = reinterpret_cast<const double2*>(recv);
  double2* out2 = reinterpret_cast<double2*>(gstrip) + goff * 3;
  unsigned long long m0 = 0ull, m1 = 0ull;
  for (int64_t e = lane; e < glen * 3; e += 32) {
    const int pose = glo + (int)(e / 3);
    const int comp = (int)(e % 3);
    double2 v = make_double2(0.0, 0.0);
    if (MAXW > 0) {
#pragma unroll
      for (int s = 0; s < NW; s++)
        if (pose >= lo[s] && pose <= hi[s]) {
          const double2 u = recv2[base[s] + (int64_t)(pose - lo[s]) * 3 + comp];
          v.x += u.x; v.y += u.y;
        }
    } else {
      for (int s = 0; s < W; s++) {
        const int l = win_all[((size_t)s * Np + a) * 2], h2 = win_all[((size_t)s * Np + a) * 2 + 1];
        if (pose >= l && pose <= h2) {
          const double2 u = recv2[(own_off[(size_t)s * (n_own + 1) + i] + (pose - l)) * 3 + comp];
          v.x += u.x; v.y += u.y;
        }
      }
    }
    out2[e] = v;
    if (v.x != 0.0 || v.y != 0.0) {  // occupancy masks of the merged strip, straight from the merged values
      m0 |= 1ull << min(63, pose / group);
      if (pose >= 1) m1 |= 1ull << min(63, (pose - 1) / group);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 |= __shfl_xor_sync(0xffffffffu, m0, o);
    m1 |= __shfl_xor_sync(0xffffffffu, m1, o);
  }
  if (lane == 0) { gmask[2 * a] = m0; gmask[2 * a + 1] = m1; }
}


// ---------------------------------------------------------------------------------------------------
// Peer-memory variant of the exchange (default when every rank can map every other rank's receive buffer with CUDA
// IPC; NVLink / NVSwitch peers). The all-to-all disappears into the map-side kernel: every warp of k_pix stores its
// finished sub-strip straight into the owner's receive buffer (posted NVLink stores, 1/W of them local), at the
// offset the owner's merge expects. All ranks hold all pose windows after the window all-gather, so every rank
// computes the same [source][owner] volume matrix and from it every receive layout -- no sizes travel. The grouped
// A22 / b2 all-reduce that follows the map-side kernel is also the barrier: it completes on a rank only after
// every rank has entered it, i.e. after every rank's map-side kernel (and its remote stores) has finished. The
// next assembly's window all-gather orders the owner's merge (a read of the receive buffer) before any peer's next
// map-side kernel writes into it.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int owner_of(int64_t a, int64_t Np, int W) { return (int)(((a + 1) * W - 1) / Np); }

// cnt[s * W + q] = poses of the sub-strips rank s holds for the pixels rank q owns
__global__ void k_cnt_matrix(int W, int64_t Np, const int32_t* __restrict__ win_all, unsigned long long* __restrict__ cnt) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = a < Np;
  const int q = valid ? owner_of(a, Np, W) : -1;
  const int q0 = __shfl_sync(0xffffffffu, q, 0);
  const bool uniform = __all_sync(0xffffffffu, q == q0);
  for (int s = 0; s < W; s++) {
    long long len = 0;
    if (valid) {
      const int lo = win_all[((size_t)s * Np + a) * 2], hi = win_all[((size_t)s * Np + a) * 2 + 1];
      len = hi >= lo ? (long long)(hi - lo + 1) : 0;
    }
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
      if ((threadIdx.x & 31) == 0 && len) atomicAdd(cnt + (size_t)s * W + q0, (unsigned long long)len);
    } else if (valid && len) {
      atomicAdd(cnt + (size_t)s * W + q, (unsigned long long)len);
    }
  }
}

struct PeerTab {
  double* recv[Handle::kPeerMax];   // receive buffer of owner q as mapped on this rank
  int64_t base[Handle::kPeerMax];   // poses in front of this rank's chunk in owner q's buffer
};

// destination of every local sub-strip: owner's buffer + my chunk's base + the strip's offset inside my chunk
// (local strips are laid out in pixel order, so inside a chunk the offsets are the local ones, shifted)
__global__ void k_strip_dst(int W, int64_t Np, const int64_t* __restrict__ stripoff, PeerTab tab, int64_t* __restrict__ dst) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  const int q = owner_of(a, Np, W);
  const int64_t a0 = Np * q / W;
  double* p = tab.recv[q] + (tab.base[q] + stripoff[a] - stripoff[a0]) * 6;
  dst[a] = (int64_t)(uintptr_t)p;
}

// the few numbers the host needs for the exchange: receive counts per source rank, my strip offsets at the
// ownership boundaries, merged strip total
__global__ void k_gather_meta(int W, int64_t Np, int64_t n_own, const int64_t* __restrict__ own_off,
                              const int64_t* __restrict__ stripoff, const int64_t* __restrict__ gstripoff,
                              int64_t* __restrict__ meta) {
  const int t = threadIdx.x;
  for (int s = t; s < W; s += blockDim.x)
    meta[s] = own_off[(size_t)s * (n_own + 1) + n_own] - own_off[(size_t)s * (n_own + 1)];  // poses received from s
  for (int q = t; q <= W; q += blockDim.x) meta[W + q] = stripoff[Np * q / W];
  if (t == 0) meta[2 * W + 1] = gstripoff[Np];
}

static bool peer_wanted(const Handle* h) {
  static const bool off = getenv("EMBA_XCHG_PEER") && atoi(getenv("EMBA_XCHG_PEER")) == 0;
  return !off && h->world > 1 && h->world <= Handle::kPeerMax && h->peer_mode != 0;
}

struct PeerMsg {  // what a rank publishes about its receive buffer (80 bytes)
  cudaIpcMemHandle_t handle;
  int64_t offset;  // of d_recv inside the allocation the handle names
  int64_t ok;
};
static_assert(sizeof(PeerMsg) == 80, "PeerMsg layout");

// (Re)allocates the receive buffers that have to grow (want[q] > 0: new capacity of rank q's, in doubles) and maps
// them on every rank. Every rank calls this with the same `want` (it follows from the all-gathered pose windows), so
// the collectives inside line up. On any failure anywhere, ALL ranks switch to the ncclSend/ncclRecv exchange for
// good (the decision is taken on all-gathered / all-reduced flags). The caller has synchronised the stream.
static void peer_note(const Handle* h, const char* what, cudaError_t e) {
  static const bool dbg = getenv("EMBA_PEER_DEBUG") && atoi(getenv("EMBA_PEER_DEBUG")) == 1;
  if (dbg) fprintf(stderr, "[emba_b200 rank %d] peer-memory exchange: %s: %s\n", h->rank, what, cudaGetErrorString(e));
}

static int peer_remap(Handle* h, const int64_t* want) {
  NcclApi* api = nccl_api();
  const int W = h->world, r = h->rank;
  const int words = (int)(sizeof(PeerMsg) / sizeof(int32_t));
  if (!h->d_peerx) EMBA_TRY(dev_alloc(h, &h->d_peerx, (int64_t)(Handle::kPeerMax + 2) * 10));
  for (int q = 0; q < W; q++)
    if (q != r && want[q] > 0 && h->peer_base[q]) {  // nothing of mine is using the old mapping any more
      cudaIpcCloseMemHandle(h->peer_base[q]);
      cudaGetLastError();
      h->peer_base[q] = nullptr;
      h->peer_recv[q] = nullptr;
    }
  PeerMsg* msg = reinterpret_cast<PeerMsg*>(h->h_pin + 600);
  PeerMsg* all = reinterpret_cast<PeerMsg*>(h->h_pin + 620);
  std::memset(msg, 0, sizeof(PeerMsg));
  msg->ok = 1;
  double* old = nullptr;
  if (want[r] > 0) {
    double* fresh = nullptr;
    if (cudaError_t e = cudaMalloc((void**)&fresh, sizeof(double) * (size_t)want[r])) { peer_note(h, "cudaMalloc", e); cudaGetLastError(); msg->ok = 0; }
    else { old = h->d_recv; h->d_recv = fresh; h->recv_cap = want[r]; }
  }
  if (msg->ok && h->d_recv) {
    if (cudaError_t e = cudaIpcGetMemHandle(&msg->handle, h->d_recv)) { peer_note(h, "cudaIpcGetMemHandle", e); cudaGetLastError(); msg->ok = 0; }
    else {
      typedef CUresult (*fn_range)(CUdeviceptr*, size_t*, CUdeviceptr);
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qr;
      CUdeviceptr base = 0;
      size_t size = 0;
      if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn ||
          ((fn_range)fn)(&base, &size, (CUdeviceptr)(uintptr_t)h->d_recv) != CUDA_SUCCESS) {
        peer_note(h, "cuMemGetAddressRange", cudaErrorUnknown); cudaGetLastError(); msg->ok = 0;
      }
      else msg->offset = (int64_t)((uintptr_t)h->d_recv - (uintptr_t)base);
    }
  }
  int32_t* stage = reinterpret_cast<int32_t*>(h->d_peerx);
  EMBA_CUDA(cudaMemcpyAsync(stage + (size_t)W * words, msg, sizeof(PeerMsg), cudaMemcpyHostToDevice, h->stream));
  if (api->allgather(stage + (size_t)W * words, stage, (size_t)words, 2, h->nccl_comm, h->stream) != 0) {
    h->err = "ncclAllGather (receive-buffer handles) failed"; return EMBA_E_NCCL;
  }
  h->launches++;
  EMBA_CUDA(cudaMemcpyAsync(all, stage, sizeof(PeerMsg) * W, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  // every rank unmapped the buffers that are being replaced BEFORE it contributed to that all-gather
  if (old) cudaFree(old);
  bool ok = true;
  for (int q = 0; q < W; q++) {
    ok = ok && all[q].ok == 1;
    if (all[q].ok != 1) peer_note(h, "a rank could not publish its receive buffer", cudaErrorUnknown);
  }
  int32_t mine = ok ? 1 : 0;
  if (ok) {
    for (int q = 0; q < W && mine; q++) {
      if (q == r || want[q] <= 0) continue;
      void* base = nullptr;
      if (cudaError_t e = cudaIpcOpenMemHandle(&base, all[q].handle, cudaIpcMemLazyEnablePeerAccess)) { peer_note(h, "cudaIpcOpenMemHandle", e); cudaGetLastError(); mine = 0; break; }
      h->peer_base[q] = base;
      h->peer_recv[q] = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + all[q].offset);
    }
    int32_t* hflag = reinterpret_cast<int32_t*>(h->h_pin + 600);
    *hflag = mine;
    EMBA_CUDA(cudaMemcpyAsync(stage, hflag, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (api->allreduce(stage, stage, 1, 2, 3, h->nccl_comm, h->stream) != 0) { h->err = "ncclAllReduce (peer mapping) failed"; return EMBA_E_NCCL; }
    h->launches++;
    EMBA_CUDA(cudaMemcpyAsync(hflag, stage, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
    ok = *hflag == 1;
    if (!ok) peer_note(h, "a rank could not map a peer's receive buffer", cudaErrorUnknown);
  }
  if (!ok) {
    for (int q = 0; q < W; q++) {
      if (h->peer_base[q]) { cudaIpcCloseMemHandle(h->peer_base[q]); cudaGetLastError(); }
      h->peer_base[q] = nullptr; h->peer_recv[q] = nullptr; h->peer_cap[q] = 0;
    }
    h->peer_mode = 0;
    return EMBA_OK;
  }
  for (int q = 0; q < W; q++) if (want[q] > 0) h->peer_cap[q] = want[q];
  h->peer_recv[r] = h->d_recv;
  h->peer_mode = 1;
  peer_note(h, "receive buffers mapped", cudaSuccess);
  return EMBA_OK;
}

// The exchange in three parts, so that the transfers overlap the map-side kernel (assemble.cu):
//   comm_exchange_prepare : (before k_pix; the pose windows are final once k_asm_pose is done) window all-gather,
//                           sub-strip lengths / offsets / merged windows, the few numbers the host needs
//   comm_exchange_step    : step s of a ring-shift all-to-all on the communication stream: my chunk for rank r+s
//                           leaves as soon as k_pix has finished that owner's pixel range, the chunk of rank r-s
//                           for me arrives
//   comm_exchange_finish  : A22 / b2 all-reduce, deterministic merge of the received sub-strips
int comm_exchange_prepare(Handle* h) {
  NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) { h->err = "multi-GPU shard without a communicator: call emba_comm_init"; return EMBA_E_NCCL; }
  const int W = h->world, r = h->rank;
  const int64_t Np = h->Np;
  const int T = 256;
  auto own0 = [&](int q) { return Np * q / W; };
  const int64_t a0 = own0(r), n_own = own0(r + 1) - a0;
  if (2 * W + 2 > 900) { h->err = "world size too large"; return EMBA_E_SUPPORT; }
  EMBA_TRY(dev_reserve(h, &h->d_win_all, &h->win_all_cap, (int64_t)W * Np * 2 + 2 * Np));
  EMBA_TRY(dev_reserve(h, &h->d_own_len, &h->own_cap, (int64_t)2 * W * (n_own + 1) + 8 * W + 32 + (int64_t)W * W));
  if (!h->d_win2) {
    EMBA_TRY(dev_alloc(h, &h->d_win2, 2 * (h->P + 1)));
    EMBA_TRY(dev_alloc(h, &h->d_gwinlo, h->P + 1));
    EMBA_TRY(dev_alloc(h, &h->d_gwinhi, h->P + 1));
    EMBA_TRY(dev_alloc(h, &h->d_gstripoff, h->P + 2));
    EMBA_TRY(dev_alloc(h, &h->d_glen, h->P + 2));
  }
  int64_t* own_len = h->d_own_len;
  int64_t* own_off = own_len + (size_t)W * (n_own + 1);
  int64_t* meta_dev = own_off + (size_t)W * (n_own + 1) + (W + 1);
  k_pack_win<<<ceil_div64(Np, T), T, 0, h->stream>>>(Np, h->d_winlo, h->d_winhi, h->d_win2);
  EMBA_LAUNCH_CHECK();
  // ncclInt32 = 2
  if (api->allgather(h->d_win2, h->d_win_all, (size_t)Np * 2, 2, h->nccl_comm, h->stream) != 0) {
    h->err = "ncclAllGather (pose windows) failed"; return EMBA_E_NCCL;
  }
  h->launches++;
  k_own_len<<<ceil_div64(Np + 1, T), T, 0, h->stream>>>(W, Np, a0, n_own, h->d_win_all, own_len, h->d_gwinlo,
                                                        h->d_gwinhi, h->d_glen);
  EMBA_LAUNCH_CHECK();
  k_zero_tail<<<1, 64, 0, h->stream>>>(W, n_own, own_len);
  EMBA_LAUNCH_CHECK();
  // ONE scan over the flattened [source][pixel] lengths (each source's tail entry is 0): own_off[s][i] is then the
  // offset of pixel i's sub-strip from source s in the receive buffer, chunk base included
  EMBA_TRY(scan_exclusive<int64_t>(h, h->stream, own_len, own_off, (int64_t)W * (n_own + 1), h->d_scan_tmp));
  EMBA_TRY(scan_exclusive<int64_t>(h, h->stream, h->d_glen, h->d_gstripoff, Np + 1, h->d_scan_tmp));
  // host needs: recv counts (W), my local strip offsets at the ownership boundaries (W+1), merged total (1):
  // gathered on the device, read back with ONE copy into pinned memory (the caller synchronises)
  k_gather_meta<<<1, 64, 0, h->stream>>>(W, Np, n_own, own_off, h->d_stripoff, h->d_gstripoff, meta_dev);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 32, meta_dev, sizeof(int64_t) * (2 * W + 2), cudaMemcpyDeviceToHost, h->stream));
  if (peer_wanted(h)) {
    // [source][owner] volumes, identical on every rank: the receive layout of every owner follows from them
    int64_t* cnt_dev = meta_dev + (2 * W + 2);
    EMBA_CUDA(cudaMemsetAsync(cnt_dev, 0, sizeof(int64_t) * W * W, h->stream));
    k_cnt_matrix<<<ceil_div64(Np, T), T, 0, h->stream>>>(W, Np, h->d_win_all, reinterpret_cast<unsigned long long*>(cnt_dev));
    EMBA_LAUNCH_CHECK();
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 256, cnt_dev, sizeof(int64_t) * W * W, cudaMemcpyDeviceToHost, h->stream));
  }
  return EMBA_OK;
}

// after the caller's synchronisation: sizes the receive / merged buffers
int comm_exchange_sizes(Handle* h) {
  const int W = h->world;
  const int64_t* meta = h->h_pin + 32;
  h->x_recv_cnt.assign(meta, meta + W);
  h->x_send_off.assign(meta + W, meta + 2 * W + 1);
  h->x_gtot = meta[2 * W + 1];
  h->x_recvbase.assign(W + 1, 0);
  for (int s = 0; s < W; s++) h->x_recvbase[s + 1] = h->x_recvbase[s] + h->x_recv_cnt[s];
  EMBA_TRY(dev_reserve(h, &h->d_gstrip, &h->gstrip_cap, h->x_gtot * 6 + h->x_gtot * 3));
  h->peer_now = false;
  if (peer_wanted(h)) {
    const int64_t* cm = h->h_pin + 256;
    h->x_cnt.assign(cm, cm + (size_t)W * W);
    int64_t want[Handle::kPeerMax] = {};
    bool any = false;
    for (int q = 0; q < W; q++) {
      int64_t need = 0;
      for (int s = 0; s < W; s++) need += h->x_cnt[(size_t)s * W + q];
      need *= 6;
      if (need > h->peer_cap[q]) { want[q] = need + need / 2 + 1024; any = true; }  // +50 %: windows drift between iterations
    }
    if (any) EMBA_TRY(peer_remap(h, want));
    h->peer_now = h->peer_mode == 1;
  }
  if (!h->peer_now) EMBA_TRY(dev_reserve(h, &h->d_recv, &h->recv_cap, h->x_recvbase[W] * 6 + h->x_recvbase[W] * 3));
  return EMBA_OK;
}

// peer mode: per-pixel destinations of this assembly's sub-strips (before the map-side kernel)
int comm_exchange_dst(Handle* h) {
  const int W = h->world, r = h->rank;
  const int64_t Np = h->Np;
  EMBA_TRY(dev_reserve(h, &h->d_dst, &h->dst_cap, Np + 1));
  PeerTab tab;
  for (int q = 0; q < Handle::kPeerMax; q++) { tab.recv[q] = nullptr; tab.base[q] = 0; }
  for (int q = 0; q < W; q++) {
    tab.recv[q] = q == r ? h->d_recv : h->peer_recv[q];
    int64_t b = 0;
    for (int s = 0; s < r; s++) b += h->x_cnt[(size_t)s * W + q];
    tab.base[q] = b;
  }
  k_strip_dst<<<ceil_div64(Np, 256), 256, 0, h->stream>>>(W, Np, h->d_stripoff, tab, h->d_dst);
  EMBA_LAUNCH_CHECK();
  return EMBA_OK;
}

// step s = 1 .. W-1 on stream st: send my sub-strips of rank (r + s)'s pixels, receive rank (r - s)'s for mine
int comm_exchange_step(Handle* h, int step, cudaStream_t st) {
  NcclApi* api = nccl_api();
  const int W = h->world, r = h->rank;
  const int to = (r + step) % W, from = (r - step + W) % W;
  const int64_t scount = (h->x_send_off[to + 1] - h->x_send_off[to]) * 6;
  const int64_t rcount = h->x_recv_cnt[from] * 6;
  if (api->group_start() != 0) { h->err = "ncclGroupStart failed"; return EMBA_E_NCCL; }
  int rc = 0;  // ncclFloat64 = 8
  if (scount > 0) rc |= api->send(h->d_strip + h->x_send_off[to] * 6, (size_t)scount, 8, to, h->nccl_comm, st);
  if (rcount > 0) rc |= api->recv(h->d_recv + h->x_recvbase[from] * 6, (size_t)rcount, 8, from, h->nccl_comm, st);
  if (api->group_end() != 0 || rc != 0) { h->err = "ncclSend/ncclRecv failed"; return EMBA_E_NCCL; }
  h->launches++;
  return EMBA_OK;
}

// all steps as ONE NCCL group (one launch): the unpipelined exchange, after the map-side kernel has finished. The
// two small all-reduces of the map blocks ride in the same group (with_a22).
int comm_exchange_all(Handle* h, cudaStream_t st, bool with_a22) {
  NcclApi* api = nccl_api();
  const int W = h->world, r = h->rank;
  if (api->group_start() != 0) { h->err = "ncclGroupStart failed"; return EMBA_E_NCCL; }
  int rc = 0;
  if (with_a22) {
    rc |= api->allreduce(h->d_A22, h->d_A22, (size_t)(3 * h->Np), 8, 0, h->nccl_comm, st);
    rc |= api->allreduce(h->d_b2, h->d_b2, (size_t)(2 * h->Np), 8, 0, h->nccl_comm, st);
  }
  for (int q = 0; q < W && !h->peer_now; q++) {
    if (q == r) continue;
    const int64_t scount = (h->x_send_off[q + 1] - h->x_send_off[q]) * 6;
    const int64_t rcount = h->x_recv_cnt[q] * 6;
    if (scount > 0) rc |= api->send(h->d_strip + h->x_send_off[q] * 6, (size_t)scount, 8, q, h->nccl_comm, st);
    if (rcount > 0) rc |= api->recv(h->d_recv + h->x_recvbase[q] * 6, (size_t)rcount, 8, q, h->nccl_comm, st);
  }
  if (api->group_end() != 0 || rc != 0) { h->err = "ncclSend/ncclRecv failed"; return EMBA_E_NCCL; }
  if (with_a22 || !h->peer_now) h->launches++;
  return EMBA_OK;
}

int comm_exchange_finish(Handle* h, bool a22_done) {
  NcclApi* api = nccl_api();
  const int W = h->world, r = h->rank;
  const int64_t Np = h->Np;
  const int T = 256;
  const int64_t a0 = Np * r / W, n_own = Np * (r + 1) / W - a0;
  int64_t* own_off = h->d_own_len + (size_t)W * (n_own + 1);
  // my own sub-strips do not travel (peer mode: the map-side kernel already stored them in place)
  if (h->x_recv_cnt[r] > 0 && !h->peer_now)
    EMBA_CUDA(cudaMemcpyAsync(h->d_recv + h->x_recvbase[r] * 6, h->d_strip + h->x_send_off[r] * 6,
                              sizeof(double) * h->x_recv_cnt[r] * 6, cudaMemcpyDeviceToDevice, h->stream));
  if (!a22_done) {
    // one NCCL group (= one launch) for the two small all-reduces (ncclFloat64 = 8, ncclSum = 0)
    if (api->group_start() != 0) { h->err = "ncclGroupStart failed"; return EMBA_E_NCCL; }
    int rc = api->allreduce(h->d_A22, h->d_A22, (size_t)(3 * Np), 8, 0, h->nccl_comm, h->stream);
    rc |= api->allreduce(h->d_b2, h->d_b2, (size_t)(2 * Np), 8, 0, h->nccl_comm, h->stream);
    if (api->group_end() != 0 || rc != 0) { h->err = "ncclAllReduce (A22, b2) failed"; return EMBA_E_NCCL; }
    h->launches++;
  }
  if (n_own > 0) {
    if (W <= 8)
      k_merge_strips<8><<<ceil_div64(n_own * 32, T), T, 0, h->stream>>>(W, Np, a0, n_own, h->d_win_all, own_off,
                                                                       h->d_recv, h->d_gwinlo, h->d_gstripoff, h->d_gstrip,
                                                                       h->pose_group, h->d_gmask2);
    else
      k_merge_strips<0><<<ceil_div64(n_own * 32, T), T, 0, h->stream>>>(W, Np, a0, n_own, h->d_win_all, own_off,
                                                                       h->d_recv, h->d_gwinlo, h->d_gstripoff, h->d_gstrip,
                                                                       h->pose_group, h->d_gmask2);
    EMBA_LAUNCH_CHECK();
  }
  h->sv_winlo = h->d_gwinlo; h->sv_winhi = h->d_gwinhi; h->sv_stripoff = h->d_gstripoff; h->sv_strip = h->d_gstrip;
  h->sv_gmask = h->d_gmask2;
  h->mask_min_len = -1;  // the merge wrote every owned pixel's masks
  h->sv_strip_total = h->x_gtot;
  return EMBA_OK;
}

void comm_destroy(Handle* h) {
  for (int q = 0; q < Handle::kPeerMax; q++) {
    if (h->peer_base[q]) { cudaIpcCloseMemHandle(h->peer_base[q]); cudaGetLastError(); }
    h->peer_base[q] = nullptr; h->peer_recv[q] = nullptr; h->peer_cap[q] = 0;
  }
  h->peer_mode = -1;
  h->peer_now = false;
  if (h->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->destroy(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_comm_unique_id(void* out128) {
  if (!out128) return EMBA_E_ARG;
  NcclApi* api = nccl_api();
  if (!api) return EMBA_E_NCCL;
  nccl_uid_t id;
  if (api->get_uid(&id) != 0) return EMBA_E_NCCL;
  std::memcpy(out128, &id, 128);
  return EMBA_OK;
}

int emba_comm_init(emba_handle_t hh, const void* id128, int32_t rank, int32_t world) {
  Handle* h = (Handle*)hh;
  if (!h || !id128) return EMBA_E_ARG;
  if (world < 1 || rank < 0 || rank >= world) { h->err = "emba_comm_init: bad rank/world"; return EMBA_E_ARG; }
  NcclApi* api = nccl_api();
  if (!api) { h->err = "emba_comm_init: libnccl.so.2 not found"; return EMBA_E_NCCL; }
  EMBA_CUDA(cudaSetDevice(h->device));
  comm_destroy(h);
  nccl_uid_t id;
  std::memcpy(&id, id128, 128);
  const int r = api->init_rank(&h->nccl_comm, world, id, rank);
  if (r != 0) { h->err = std::string("ncclCommInitRank: ") + (api->errstr ? api->errstr(r) : "error"); h->nccl_comm = nullptr; return EMBA_E_NCCL; }
  EMBA_TRY(emba_set_shard(hh, rank, world));
  // the first collective on a communicator sets up its channels (seconds): pay for it here, not inside the first
  // window's pre-pass
  if (world > 1) {
    EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
    EMBA_TRY(comm_allreduce(h, h->d_flags, 16, 0));
    // ... and the first all-gather sets up its own connections (2 GPUs: 0.3 s inside the first emba_set_events)
    if (!h->d_peerx) EMBA_TRY(dev_alloc(h, &h->d_peerx, (int64_t)(Handle::kPeerMax + 2) * 10));
    if (world <= 256)
      EMBA_TRY(comm_allgather_i32(h, h->d_flags, reinterpret_cast<int32_t*>(h->d_peerx), 1));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
    // peer-memory strip exchange: map every rank's receive buffer now (context creation on the peers and the IPC
    // opens take tens of milliseconds), with a starting capacity that later windows grow on demand
    if (peer_wanted(h)) {
      int64_t want[Handle::kPeerMax] = {};
      for (int q = 0; q < world; q++) want[q] = (int64_t)4 << 20;  // 32 MB
      EMBA_TRY(peer_remap(h, want));
    }
  }
  return EMBA_OK;
}

}  // extern "C"
