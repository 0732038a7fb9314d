// Device-wide primitives of the pre-pass and the assembly, hand-written for sm_100a (no CUB / Thrust anywhere in the
// library): exclusive prefix sums and a stable LSD radix sort of (key, value) pairs with 8-bit digits.
//
//   scan : 2048 elements per CTA (256 threads x 8 consecutive items through a padded shared tile, warp-shuffle scan
//          of the per-thread sums); three phases (block sums, recursive scan of the sums, local scan + offset).
//   sort : per pass  k_rs_hist (256-bin histogram per 4096-key tile, written digit-major)  ->  one exclusive scan of
//          the [digit][tile] table (global offsets of every (tile, digit) run)  ->  k_rs_scatter. The scatter keeps
//          the sort stable without any cross-warp ordering traffic: warp w owns the w-th contiguous 512-key slice of
//          the tile, ranks the 32 keys of a round with match.any (peers below me in the warp), and advances its own
//          per-digit cursor; cursors start at  table offset + keys of that digit in the warps before me.
#include <algorithm>
#include <cstring>

#include "emba_internal.cuh"

namespace emba {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

template <typename T>
__device__ __forceinline__ T block_exclusive_scan_sums(T v, T* warp_tot, T& total) {
  // exclusive scan of one value per thread over the CTA (256 threads); total = sum over the CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  T base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; w++) {
    const T t = warp_tot[w];
    if (w < warp) base += t;
    tot += t;
  }
  total = tot;
  return base + inc - v;
}

// phase 1: per-tile sums
template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_sums(const T* __restrict__ in, int64_t count, T* __restrict__ bsum) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  T s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int64_t i = base + k * kScanThreads + threadIdx.x;
    if (i < count) s += in[i];
  }
  __shared__ T sh[kScanThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) t += sh[w];
    bsum[blockIdx.x] = t;
  }
}

// phase 3 (and the single-tile case): local exclusive scan + tile offset. offs == nullptr: offset 0.
template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const T* in, T* out, int64_t count, const T* offs) {  // in may alias out
  __shared__ T tile[kScanTile + kScanTile / 32];
  __shared__ T warp_tot[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int j = k * kScanThreads + threadIdx.x;
    const int64_t i = base + j;
    tile[j + (j >> 5)] = i < count ? in[i] : (T)0;
  }
  __syncthreads();
  T v[kScanItems];
  T s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int j = threadIdx.x * kScanItems + k;
    v[k] = tile[j + (j >> 5)];
    s += v[k];
  }
  T total;
  T run = block_exclusive_scan_sums<T>(s, warp_tot, total) + (offs ? offs[blockIdx.x] : (T)0);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int j = threadIdx.x * kScanItems + k;
    tile[j + (j >> 5)] = run;
    run += v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    const int j = k * kScanThreads + threadIdx.x;
    const int64_t i = base + j;
    if (i < count) out[i] = tile[j + (j >> 5)];
  }
}

size_t scan_scratch_bytes(int64_t count) {
  // block sums of every level, 8 bytes per entry, 256-byte aligned levels
  size_t tot = 0;
  int64_t c = count;
  while (c > kScanTile) {
    c = (c + kScanTile - 1) / kScanTile;
    tot += (((size_t)c * 8 + 255) / 256) * 256;
  }
  return tot + 256;
}

template <typename T>
int scan_exclusive(Handle* h, cudaStream_t st, const T* in, T* out, int64_t count, void* scratch) {
  if (count <= 0) return EMBA_OK;
  const int64_t nb = (count + kScanTile - 1) / kScanTile;
  if (nb == 1) {
    k_scan_apply<T><<<1, kScanThreads, 0, st>>>(in, out, count, nullptr);
    EMBA_LAUNCH_CHECK();
    return EMBA_OK;
  }
  T* bsum = reinterpret_cast<T*>(scratch);
  char* next = reinterpret_cast<char*>(scratch) + (((size_t)nb * 8 + 255) / 256) * 256;
  k_scan_sums<T><<<(unsigned)nb, kScanThreads, 0, st>>>(in, count, bsum);
  EMBA_LAUNCH_CHECK();
  EMBA_TRY(scan_exclusive<T>(h, st, bsum, bsum, nb, next));  // in place: every tile reads before it writes
  k_scan_apply<T><<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, count, bsum);
  EMBA_LAUNCH_CHECK();
  return EMBA_OK;
}
template int scan_exclusive<int32_t>(Handle*, cudaStream_t, const int32_t*, int32_t*, int64_t, void*);
template int scan_exclusive<int64_t>(Handle*, cudaStream_t, const int64_t*, int64_t*, int64_t, void*);
template int scan_exclusive<uint32_t>(Handle*, cudaStream_t, const uint32_t*, uint32_t*, int64_t, void*);

// ---------------------------------------------------------------------------------------------------
// radix sort
// ---------------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsRounds = 16;                       // keys per thread
constexpr int kRsTile = kRsThreads * kRsRounds;     // 4096 keys per CTA
constexpr int kRsWarpKeys = 32 * kRsRounds;         // 512 contiguous keys per warp

__global__ void __launch_bounds__(kRsThreads)
k_rs_hist(const uint32_t* __restrict__ keys, int64_t count, int shift, int64_t ntiles, uint32_t* __restrict__ table) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int r = 0; r < kRsRounds; r++) {
    const int64_t i = base + r * kRsThreads + threadIdx.x;
    if (i < count) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  table[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// vals_in == nullptr: the values are the key indices (first pass of an argsort)
__global__ void __launch_bounds__(kRsThreads)
k_rs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals_in, int64_t count, int shift,
             int64_t ntiles, const uint32_t* __restrict__ table, uint32_t* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t cnt[kRsThreads / 32][256];  // per-warp digit counts, then per-warp cursors
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kRsThreads / 32) * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)warp * kRsWarpKeys;
  uint32_t k[kRsRounds];
  const uint32_t lt = (1u << lane) - 1u;
  // pass A: count
#pragma unroll
  for (int r = 0; r < kRsRounds; r++) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < count;
    k[r] = ok ? keys[i] : 0xFFFFFFFFu;
    const uint32_t d = ok ? ((k[r] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & lt) == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // cursor of (warp, digit) = global offset of the (tile, digit) run + keys of that digit in the warps before
  {
    const int d = threadIdx.x;
    uint32_t run = table[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsThreads / 32; w++) {
      const uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass B: place
#pragma unroll
  for (int r = 0; r < kRsRounds; r++) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < count;
    const uint32_t d = ok ? ((k[r] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (ok) {
      const uint32_t pos = cnt[warp][d] + __popc(peers & lt);
      if (keys_out) keys_out[pos] = k[r];
      vals_out[pos] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    __syncwarp();
    if (ok && (peers & lt) == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
}

size_t radix_scratch_bytes(int64_t count) {
  const int64_t ntiles = (count + kRsTile - 1) / kRsTile;
  const size_t table = (((size_t)ntiles * 256 * 4 + 255) / 256) * 256;
  return table + scan_scratch_bytes(ntiles * 256);
}

// Stable sort of (key, value) pairs by key bits [0, nbits). Buffers ping-pong between (k0, v0) and (k1, v1), all four
// are overwritten when there are two passes or more; iota == true means "values = indices" (v0 need not be
// initialised). Returns through *which the buffer holding the result (0 or 1). keep_keys == false skips writing the
// keys of the last pass (the caller only wants the permutation).
int radix_sort_pairs(Handle* h, cudaStream_t st, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, int64_t count,
                     int nbits, void* scratch, bool iota, bool keep_keys, int* which) {
  *which = 0;
  if (count <= 0) return EMBA_OK;
  const int64_t ntiles = (count + kRsTile - 1) / kRsTile;
  uint32_t* table = reinterpret_cast<uint32_t*>(scratch);
  void* scan_scr = reinterpret_cast<char*>(scratch) + (((size_t)ntiles * 256 * 4 + 255) / 256) * 256;
  const int passes = std::max(1, (nbits + 7) / 8);
  uint32_t* kb[2] = {k0, k1};
  uint32_t* vb[2] = {v0, v1};
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    const int shift = 8 * p;
    k_rs_hist<<<(unsigned)ntiles, kRsThreads, 0, st>>>(kb[cur], count, shift, ntiles, table);
    EMBA_LAUNCH_CHECK();
    EMBA_TRY(scan_exclusive<uint32_t>(h, st, table, table, ntiles * 256, scan_scr));
    const bool last = p == passes - 1;
    k_rs_scatter<<<(unsigned)ntiles, kRsThreads, 0, st>>>(kb[cur], (p == 0 && iota) ? nullptr : vb[cur], count, shift,
                                                         ntiles, table, (last && !keep_keys) ? nullptr : kb[cur ^ 1],
                                                         vb[cur ^ 1]);
    EMBA_LAUNCH_CHECK();
    cur ^= 1;
  }
  *which = cur;
  return EMBA_OK;
}

// ---------------------------------------------------------------------------------------------------
// host <-> device copies of caller memory
// ---------------------------------------------------------------------------------------------------
bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// both pinned stages and their events, created together on first use
cudaError_t uploader_init(Uploader& up) {
  cudaError_t e;
  for (int i = 0; i < 2; i++) {
    if (!up.stage[i]) {
      if ((e = cudaMallocHost(&up.stage[i], kStageBytes)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&up.ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

// returns a cudaError_t; every copy is checked
cudaError_t upload_bytes(Uploader& up, cudaStream_t st, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  if (is_pinned(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  cudaError_t e;
  for (int i = 0; i < 2; i++) {
    if (!up.stage[i]) {
      if ((e = cudaMallocHost(&up.stage[i], kStageBytes)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&up.ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
  }
  for (size_t off = 0; off < bytes; off += kStageBytes) {
    const size_t len = std::min(kStageBytes, bytes - off);
    const int b = up.turn;
    up.turn ^= 1;
    if ((e = cudaEventSynchronize(up.ev[b])) != cudaSuccess) return e;  // the DMA that last read this buffer
    std::memcpy(up.stage[b], (const char*)src + off, len);
    if ((e = cudaMemcpyAsync((char*)dst + off, up.stage[b], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(up.ev[b], st)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}
void uploader_free(Uploader& up) {
  for (int i = 0; i < 2; i++) {
    if (up.stage[i]) cudaFreeHost(up.stage[i]);
    if (up.ev[i]) cudaEventDestroy(up.ev[i]);
    up.stage[i] = nullptr; up.ev[i] = nullptr;
  }
}


// device -> host into caller memory: direct when the destination is page-locked, otherwise the DMA of chunk k+1 into
// one pinned stage overlaps the memcpy of chunk k out of the other. Returns after the last byte has landed.
cudaError_t download_bytes(Uploader& up, cudaStream_t st, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  cudaError_t e;
  if (is_pinned(dst)) {
    if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
  }
  for (int i = 0; i < 2; i++) {
    if (!up.stage[i]) {
      if ((e = cudaMallocHost(&up.stage[i], kStageBytes)) != cudaSuccess) return e;
      if ((e = cudaEventCreateWithFlags(&up.ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
  }
  size_t prev_off = 0, prev_len = 0;
  int prev_b = -1;
  for (size_t off = 0; off < bytes; off += kStageBytes) {
    const size_t len = std::min(kStageBytes, bytes - off);
    const int b = up.turn;
    up.turn ^= 1;
    if ((e = cudaEventSynchronize(up.ev[b])) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(up.stage[b], (const char*)src + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(up.ev[b], st)) != cudaSuccess) return e;
    if (prev_b >= 0) {
      if ((e = cudaEventSynchronize(up.ev[prev_b])) != cudaSuccess) return e;
      std::memcpy((char*)dst + prev_off, up.stage[prev_b], prev_len);
    }
    prev_b = b; prev_off = off; prev_len = len;
  }
  if ((e = cudaEventSynchronize(up.ev[prev_b])) != cudaSuccess) return e;
  std::memcpy((char*)dst + prev_off, up.stage[prev_b], prev_len);
  return cudaSuccess;
}

}  // namespace emba
