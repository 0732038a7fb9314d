// LEGM::formNormalEq / formNormalEqIRLS + applyL2Reg on the device (reference src/emba/model.cpp:316-719).
//
//   1. active set: pixels with num_ev_map >= thres in ascending index (the std::set order) -> mask + scan
//   2. k_asm_pose: one CTA per work item (a run of measurements touching the same control-pose pair, static
//      per window). Each thread forms one Jacobian row [Jc(6) Jp(6) | dp(2) | e]; the 13x13 outer products
//      [Jc Jp e]^T [Jc Jp e] (A11 blocks, b1) are accumulated in registers from a shared-memory tile and
//      reduced with warp shuffles; per-item partials are combined in a fixed order (deterministic).
//      The row is also written out (128-byte record) for the map side.
//   3. map side: rows are grouped by target pixel WITHOUT a global sort. The evaluation's histogram is exactly the
//      segment-length table (scan -> segment offsets), and the value its counting atomic returned is a unique slot
//      of the row inside its pixel's segment: k_place drops every row id into place (one 4-byte scattered store per
//      row), k_seg_sort then orders each segment by row id (bitonic network in registers, one warp per pixel; CTA-wide
//      shared-memory network for the rare segments above 1024 rows), which makes the summation order canonical
//      (= stable sort by pixel) again. Both need only the evaluation, so they run on a side stream beside step 2;
//      one warp per active pixel then reduces its rows in that fixed order into A22 / b2 and the pixel's A12 strip
//      (3x2 block per control pose in the pixel's pose window) -- a deterministic segmented reduction.
#include <climits>
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "emba_internal.cuh"

namespace emba {

PanoCam make_cam(const Handle* h);
int comm_allreduce(Handle* h, void* buf, int64_t count, int dtype);
int comm_exchange_prepare(Handle* h);
int comm_exchange_sizes(Handle* h);
int comm_exchange_dst(Handle* h);
int comm_exchange_step(Handle* h, int step, cudaStream_t st);
int comm_exchange_all(Handle* h, cudaStream_t st, bool with_a22);
int comm_exchange_finish(Handle* h, bool a22_done);

// ---------------------------------------------------------------------------------------------------
// flag = pixel active (global count >= thres, model.cpp:333); segcnt = this rank's rows on it if active, else 0
__global__ void k_active_flags(const int32_t* __restrict__ hist, const int32_t* __restrict__ hist_loc, int64_t P,
                               int thres, int32_t* __restrict__ flag, int32_t* __restrict__ segcnt) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int f = hist[p] >= thres ? 1 : 0;
  flag[p] = f;
  segcnt[p] = f ? ((hist_loc[p] + 7) & ~7) : 0;  // segments start on 32-byte boundaries (vector loads of the segment sort)
}

// aidx / rowbase: exclusive scans of flag / segcnt. rowbase is rewritten in place to "first row of the pixel's
// segment, or -1 if the pixel is inactive" (what k_place looks up per row).
__global__ void k_active_fill(const int32_t* __restrict__ flag, const int32_t* __restrict__ aidx,
                              const int32_t* __restrict__ hist_loc, int32_t* __restrict__ rowbase, int64_t P,
                              int32_t* __restrict__ amap, int32_t* __restrict__ apix, int2* __restrict__ win,
                              double4* __restrict__ H3, int32_t* __restrict__ segoff, int32_t* __restrict__ segend,
                              int64_t* __restrict__ totals) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < P;
  // the active index also rides in the spare lane of the Hessian entry, so the assembly kernel gets it with the
  // gather it does anyway
  const int f = valid ? flag[p] : 0;
  const int32_t a = valid ? aidx[p] : 0;
  const int32_t base = valid ? rowbase[p] : 0;
  const int32_t cnt = f ? hist_loc[p] : 0;
  // rows on active pixels (this rank): one atomic per warp
  {
    long long cntw = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntw += __shfl_xor_sync(0xffffffffu, cntw, o);
    if ((threadIdx.x & 31) == 0 && cntw) atomicAdd(reinterpret_cast<unsigned long long*>(totals + 1), (unsigned long long)cntw);
  }
  if (!valid) return;
  reinterpret_cast<double*>(H3 + p)[3] = __longlong_as_double((long long)(f ? a : -1));
  if (p == P - 1) totals[0] = (int64_t)a + f;
  if (f) {
    amap[p] = a;
    apix[a] = (int32_t)p;
    win[a] = make_int2(INT_MAX, -1);
    segoff[a] = base;
    segend[a] = base + cnt;
  } else {
    amap[p] = -1;
    rowbase[p] = -1;
  }
}

// every inlier row on an active pixel goes to its slot of the pixel's segment (a separate pass: fused into the
// pose-side kernel the scattered 4-byte stores cost that HBM-bound kernel 2.8 ms on C4, alone they take 1.7 ms)
// ROWS rows per thread: the dependent gather chain pix -> rowbase[pix] -> store of all of them is in flight at once
// (the one-row version issued 7 % of the time: pure latency). CG: the table gather and the scattered store bypass L1
// (ld.global.cg / st.global.cg) -- neither has any reuse there.
template <int ROWS, bool CG>
__global__ void __launch_bounds__(256)
k_place(int64_t Mc, const int32_t* __restrict__ pix, const int32_t* __restrict__ slot,
        const int32_t* __restrict__ rowbase, uint32_t* __restrict__ sval) {
  const int64_t m0 = (int64_t)blockIdx.x * (256 * ROWS) + threadIdx.x;
  int32_t p[ROWS], sl[ROWS], base[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    const int64_t m = m0 + k * 256;
    p[k] = m < Mc ? __ldcs(pix + m) : -1;
    sl[k] = m < Mc ? __ldcs(slot + m) : 0;
  }
#pragma unroll
  for (int k = 0; k < ROWS; k++) base[k] = p[k] >= 0 ? (CG ? __ldcg(rowbase + p[k]) : rowbase[p[k]]) : -1;
#pragma unroll
  for (int k = 0; k < ROWS; k++)
    if (base[k] >= 0) {
      if (CG) __stcg(sval + (int64_t)base[k] + sl[k], (uint32_t)(m0 + k * 256));
      else sval[(int64_t)base[k] + sl[k]] = (uint32_t)(m0 + k * 256);
    }
}

// ---------------------------------------------------------------------------------------------------
// Segment sort: ascending row ids inside every pixel's segment. "Normalised" bitonic network (every
// compare-exchange puts the minimum at the lower index: first step of a merge of size k pairs i with i ^ (k-1), the
// following steps i with i ^ j), so a segment padded with +inf behind its last element sorts correctly.
// ---------------------------------------------------------------------------------------------------
constexpr int kSegWarps = 4;
constexpr int kSegRegCap = 1024;   // longest segment a single warp sorts in registers (32 per lane)

// element i = lane * K + r  (blocked: a lane's K keys are consecutive)
// KMIN = 2: full sort. KMIN = 32 K: only the last merge (the two halves are already sorted).
// Code size matters here: fully unrolled, the six networks (K = 1 .. 32) are ~150 KB of straight-line SASS that every
// warp streams through once per segment, and the kernel spent 17.8 of its 23.3 stall cycles per issue waiting for
// instructions (ncu "no_instruction"). So only the in-register steps (compile-time register indices) are unrolled; the
// merge sizes above K and their cross-lane steps are runtime loops whose bodies -- one exchange step over K registers
// and the in-register tail -- are reused by every iteration.
template <int K>
__device__ __forceinline__ void bitonic_in_regs(uint32_t (&a)[K], int jtop) {  // steps j = jtop, jtop/2, .. 1 (jtop < K)
#pragma unroll
  for (int j = K / 2; j >= 1; j >>= 1) {
    if (j <= jtop) {
#pragma unroll
      for (int r = 0; r < K; r++) {
        const int q = r ^ j;
        if (q > r) { const uint32_t lo = min(a[r], a[q]), hi = max(a[r], a[q]); a[r] = lo; a[q] = hi; }
      }
    }
  }
}

template <int K, int KMIN = 2>
__device__ __forceinline__ void warp_bitonic(uint32_t (&a)[K], int lane) {
  // merges that stay inside a lane: compile-time network
#pragma unroll
  for (int k = KMIN; k <= K; k <<= 1) {
#pragma unroll
    for (int r = 0; r < K; r++) {  // flip step: i <-> i ^ (k - 1)
      const int q = r ^ (k - 1);
      if (q > r) { const uint32_t lo = min(a[r], a[q]), hi = max(a[r], a[q]); a[r] = lo; a[q] = hi; }
    }
#pragma unroll
    for (int j = k >> 2; j >= 1; j >>= 1) {
#pragma unroll
      for (int r = 0; r < K; r++) {
        const int q = r ^ j;
        if (q > r) { const uint32_t lo = min(a[r], a[q]), hi = max(a[r], a[q]); a[r] = lo; a[q] = hi; }
      }
    }
  }
  // merges across lanes: runtime loops over the merge size and the cross-lane steps
  constexpr int k0 = (KMIN > 2 * K) ? KMIN : 2 * K;
#pragma unroll 1
  for (int k = k0; k <= 32 * K; k <<= 1) {
    {
      const int lm = k / K - 1;  // flip step: lane ^= lm, r -> K - 1 - r
      const bool lower = (lane & ((lm + 1) >> 1)) == 0;  // the top flipped lane bit decides who is the lower index
      if (K == 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, a[0], lm);
        a[0] = lower ? min(a[0], o) : max(a[0], o);
      } else {
#pragma unroll
        for (int r = 0; r < K / 2; r++) {  // my a[r] meets the partner's a[K-1-r] and vice versa
          const uint32_t ox = __shfl_xor_sync(0xffffffffu, a[K - 1 - r], lm);
          const uint32_t oy = __shfl_xor_sync(0xffffffffu, a[r], lm);
          a[r] = lower ? min(a[r], ox) : max(a[r], ox);
          a[K - 1 - r] = lower ? min(a[K - 1 - r], oy) : max(a[K - 1 - r], oy);
        }
      }
    }
#pragma unroll 1
    for (int j = k >> 2; j >= K; j >>= 1) {
      const int lm = j / K;
      const bool lower = (lane & lm) == 0;
#pragma unroll
      for (int r = 0; r < K; r++) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, a[r], lm);
        a[r] = lower ? min(a[r], o) : max(a[r], o);
      }
    }
    if (K > 1) bitonic_in_regs<K>(a, K / 2);
  }
}

// (An adaptive front end -- a few odd-even transposition passes while the segment is unsorted -- was measured and
// dropped: rows reach their segment through atomics of ~300 k measurements in flight, and not one C4 segment in 350 k
// was sorted after 6 passes.)
// a lane's K consecutive keys travel as 32-byte (K >= 8), 16-byte (K = 4) or 8-byte (K = 2) vectors: segments start
// on 32-byte boundaries, so a K-key group costs K/8 load instructions instead of K scalar ones that each touch 32
// sectors (the scalar version was LSU-bound). Reads may run past the segment's end (masked; the buffer has a tail
// pad), stores never do.
template <int K, int KMIN = 2>
__device__ __forceinline__ void seg_sort_regs(uint32_t* __restrict__ seg, int L, int lane) {
  uint32_t a[K];
  const int i0 = lane * K;
  if (K >= 8) {
#pragma unroll
    for (int v = 0; v < K / 8; v++) {
      const uint4 lo = *reinterpret_cast<const uint4*>(seg + i0 + 8 * v);
      const uint4 hi = *reinterpret_cast<const uint4*>(seg + i0 + 8 * v + 4);
      a[8 * v] = lo.x; a[8 * v + 1] = lo.y; a[8 * v + 2] = lo.z; a[8 * v + 3] = lo.w;
      a[8 * v + 4] = hi.x; a[8 * v + 5] = hi.y; a[8 * v + 6] = hi.z; a[8 * v + 7] = hi.w;
    }
  } else if (K == 4) {
    const uint4 q = *reinterpret_cast<const uint4*>(seg + i0);
    a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
  } else if (K == 2) {
    const uint2 q = *reinterpret_cast<const uint2*>(seg + i0);
    a[0] = q.x; a[1] = q.y;
  } else {
    a[0] = seg[i0];
  }
#pragma unroll
  for (int r = 0; r < K; r++)
    if (i0 + r >= L) a[r] = 0xFFFFFFFFu;
  warp_bitonic<K, KMIN>(a, lane);
  if (K >= 4) {
#pragma unroll
    for (int v = 0; v < K / 4; v++) {
      const int i = i0 + 4 * v;
      if (i + 4 <= L) *reinterpret_cast<uint4*>(seg + i) = make_uint4(a[4 * v], a[4 * v + 1], a[4 * v + 2], a[4 * v + 3]);
      else {
#pragma unroll
        for (int r = 0; r < 4; r++) if (i + r < L) seg[i + r] = a[4 * v + r];
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < K; r++) if (i0 + r < L) seg[i0 + r] = a[r];
  }
}

__global__ void __launch_bounds__(kSegWarps * 32, 6)
k_seg_sort(const int64_t* __restrict__ np_dev, const int32_t* __restrict__ segoff, const int32_t* __restrict__ segend,
           uint32_t* __restrict__ sval, int32_t* __restrict__ longlist) {
  const int lane = threadIdx.x & 31;
  const int64_t Np = np_dev[0];  // the active-pixel count is still on its way to the host when this is enqueued
  const int64_t nw = (int64_t)gridDim.x * kSegWarps;
  for (int64_t a = (int64_t)blockIdx.x * kSegWarps + (threadIdx.x >> 5); a < Np; a += nw) {
    const int s0 = segoff[a];
    const int L = segend[a] - s0;
    uint32_t* seg = sval + s0;
    if (L <= 1) continue;
    if (L <= 32) seg_sort_regs<1>(seg, L, lane);
    else if (L <= 64) seg_sort_regs<2>(seg, L, lane);
    else if (L <= 128) seg_sort_regs<4>(seg, L, lane);
    else if (L <= 256) seg_sort_regs<8>(seg, L, lane);
    else if (L <= 512) seg_sort_regs<16>(seg, L, lane);
    else if (L <= kSegRegCap) seg_sort_regs<32>(seg, L, lane);
    else if (lane == 0) longlist[1 + atomicAdd(&longlist[0], 1)] = (int32_t)a;  // k_seg_sort_long takes it
  }
}

// segments of 1025 .. 4096 rows (a few per cent of the pixels): still one warp each. The 1024-key chunks are sorted
// with the register network above, then merged pairwise by the LAST stage of the same network over 64 and 128 keys
// per lane (a full 2048 / 4096-key network does not unroll into registers; one merge stage does). Chunks and merged
// blocks pass through global memory (L1) because the blocked layouts of the three key counts differ.
constexpr int kSegLongWarps = 2;
constexpr int kSegLongCap = 4096;
__global__ void __launch_bounds__(kSegLongWarps * 32)
k_seg_sort_long(const int32_t* __restrict__ longlist, const int32_t* __restrict__ segoff,
                const int32_t* __restrict__ segend, uint32_t* __restrict__ sval) {
  const int lane = threadIdx.x & 31;
  const int nlong = longlist[0];
  const int nw = gridDim.x * kSegLongWarps;
  for (int li = blockIdx.x * kSegLongWarps + (threadIdx.x >> 5); li < nlong; li += nw) {
    const int a = longlist[1 + li];
    const int s0 = segoff[a];
    const int L = segend[a] - s0;
    if (L > kSegLongCap) continue;
    uint32_t* seg = sval + s0;
    for (int c = 0; c < L; c += kSegRegCap) seg_sort_regs<32>(seg + c, min(kSegRegCap, L - c), lane);
    __syncwarp();
    seg_sort_regs<64, 2048>(seg, min(2048, L), lane);
    if (L > 2048) {
      if (L > 3072) seg_sort_regs<64, 2048>(seg + 2048, L - 2048, lane);
      __syncwarp();
      seg_sort_regs<128, 4096>(seg, L, lane);
    }
    __syncwarp();
  }
}

// segments above 4096 rows (none in the benchmark workloads): one CTA each, network in shared memory (up to
// cap_smem ids) or, beyond that, in place in global memory
__global__ void __launch_bounds__(512)
k_seg_sort_huge(const int32_t* __restrict__ longlist, const int32_t* __restrict__ segoff,
                const int32_t* __restrict__ segend, uint32_t* __restrict__ sval, int cap_smem) {
  extern __shared__ uint32_t sbuf[];
  const int nlong = longlist[0];
  for (int li = blockIdx.x; li < nlong; li += gridDim.x) {
    const int a = longlist[1 + li];
    const int s0 = segoff[a];
    const int L = segend[a] - s0;
    if (L <= kSegLongCap) continue;
    uint32_t* seg = sval + s0;
    const bool in_smem = L <= cap_smem;
    uint32_t* d = in_smem ? sbuf : seg;
    __syncthreads();
    if (in_smem) {
      for (int i = threadIdx.x; i < L; i += blockDim.x) sbuf[i] = seg[i];
    }
    __syncthreads();
    int n2 = 1;
    while (n2 < L) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j >= 1; j >>= 1) {
        const int mask = (j == (k >> 1)) ? (k - 1) : j;  // flip step first, then plain steps
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
          const int q = i ^ mask;
          if (q > i && q < L) {
            const uint32_t x = d[i], y = d[q];
            if (y < x) { d[i] = y; d[q] = x; }
          }
        }
        __syncthreads();
      }
    }
    if (in_smem) {
      for (int i = threadIdx.x; i < L; i += blockDim.x) seg[i] = sbuf[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
constexpr int kAsmThreads = 256;

// D += A * B with the fp64 tensor-core shape m8n8k4 (A 8x4 row-major, B 4x8 column-major). Lane l holds
// A[l/4][l%4], B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One 256-row x 128-byte box of the Jacobian-row tensor, shared -> global through the TMA (bulk async group).
__device__ __forceinline__ void tma_store_rows(const CUtensorMap* map, const void* smem, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(0),
               "r"(row0), "r"(smem_u32(smem))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Work item -> A11 / b1 partial + Jacobian rows.
//   * every thread forms one row [Jc(6) Jp(6) e | dp(2) meta] and drops it into the shared tile as a 128-byte
//     record; the 16-byte chunks of record r sit at chunk ^ (r % 8) -- the TMA's 128-byte swizzle, which also
//     makes the row-wise writes and the fragment reads below bank-conflict free;
//   * full tiles leave for global memory as ONE TMA tensor store (the TMA undoes the swizzle), overlapping the
//     next tile's gathers; the tail tile of an item is copied by the threads;
//   * the 13 x 13 products [Jc Jp e]^T [Jc Jp e] run on the fp64 tensor cores: per warp, 32 rows = 8 k-steps of
//     m8n8k4 on the field blocks (0..7) x (0..7), (0..7) x (8..15), (8..15) x (8..15); one fragment serves as A
//     and as B. Six accumulator doubles per thread instead of a 91-entry triangle spread over thread roles.
template <int COST>
__global__ void __launch_bounds__(kAsmThreads, 3)
k_asm_pose(const __grid_constant__ CUtensorMap jmap, const WorkItem* __restrict__ items,
           const MeasRec* __restrict__ rec, const double* __restrict__ Ktab, const double4* __restrict__ RotTab,
           const double4* __restrict__ JacTab, const double2* __restrict__ G2, const double4* __restrict__ H3,
           const double2* __restrict__ dp_in, const double* __restrict__ e_in, const int32_t* __restrict__ pix_in,
           PanoCam cam, double eta, double* __restrict__ jrec, int2* __restrict__ win,
           double* __restrict__ acc_part) {
  __shared__ __align__(1024) double tile[kAsmThreads * kRecDoubles];  // [row][16], chunk-swizzled
  __shared__ double knots[2 * kKnotStride];  // knot-interval data of cp_c and cp_p: uniform over the work item
  const WorkItem it = items[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid < 2 * kKnotStride)
    knots[tid] = Ktab[(size_t)(tid < kKnotStride ? it.cp_c : it.cp_p) * kKnotStride + (tid % kKnotStride)];
  __syncthreads();
  const double* knot_c = knots;
  const double* knot_p = knots + kKnotStride;
  const int lane = tid & 31, warp = tid >> 5;
  double c00a = 0.0, c00b = 0.0, c01a = 0.0, c01b = 0.0, c11a = 0.0, c11b = 0.0;
  // fragment addressing: k-step s covers rows 32*warp + 8*(s/2) + (s%2) + 2*(lane%4); field lane/4 (+8)
  const int fq = lane >> 2, kk = lane & 3;
  const unsigned long long meta_lo = (unsigned long long)((uint32_t)it.cp_c | ((uint32_t)it.cp_p << 16));
  double2* trow = reinterpret_cast<double2*>(tile + tid * kRecDoubles);
  const int swz = tid & 7;
  bool store_pending = false;
  // streamed per-measurement inputs of the NEXT tile are requested before the tensor-core phase of the current
  // one, and all of them at once: the dependent gathers (map Hessian / gradient by pixel, batch tables by batch)
  // then sit one memory latency behind them instead of three
  int32_t pix_n = -1;
  double4 r0_n = make_double4(0.0, 0.0, 0.0, 0.0);
  double2 dp_n = make_double2(0.0, 0.0);
  double e_n = 0.0;
  auto fetch = [&](int t0) {
    const int j = t0 + tid;
    if (j < it.count) {
      const int64_t m = (int64_t)it.start + j;
      pix_n = pix_in[m];
      r0_n = ldg256(rec + m);
      dp_n = dp_in[m];
      e_n = e_in[m];
    } else {
      pix_n = -1;
    }
  };
  fetch(0);
  for (int t0 = 0; t0 < it.count; t0 += kAsmThreads) {
    const int j = t0 + tid;
    double row[13];
#pragma unroll
    for (int k = 0; k < 13; k++) row[k] = 0.0;
    double d0 = 0.0, d1 = 0.0;
    const int64_t m = (int64_t)it.start + j;
    if (j < it.count) {
      const int32_t pix = pix_n;
      double4 Hh = make_double4(0.0, 0.0, 0.0, 0.0);
      double2 g = make_double2(0.0, 0.0);
      int32_t a = -1;  // model.cpp:396-412: outliers and inactive pixels are skipped
      // (Forcing the four batch-table gathers below to be issued here, together with the map gather and before the
      // active test -- the trick that gained 10 % in k_eval -- was measured on C4: 7.15 ms with all four, 6.4 ms with
      // the two RotTab entries, against 5.24 ms as written: the extra live registers spill at 80 per thread.)
      if (pix >= 0) {
        Hh = ldg256(H3 + pix);
        g = G2[pix];
        a = (int32_t)__double_as_longlong(Hh.w);
      }
      if (a >= 0) {
        const double4 r0 = r0_n;
        const double bx = r0.x, by = r0.y, bz = r0.z;
        const unsigned long long rw = (unsigned long long)__double_as_longlong(r0.w);
        const uint32_t bc = (uint32_t)rw & 0x7FFFFFFFu, bp = (uint32_t)(rw >> 32);
        const double2 dpv = dp_n;
        double e = e_n;
        // temp = Gpm + dp^T * G2pm (model.cpp:233-238)
        const double h0 = g.x + dpv.x * Hh.x + dpv.y * Hh.y;
        const double h1 = g.y + dpv.x * Hh.y + dpv.y * Hh.z;
        double vc[3], vp[3], wc[3], wp[3];
        {
          const double4 rt = ldg256(RotTab + bc);
          const double4 jt = ldg256(JacTab + bc);
          double X, Y, Z, M[6];
          rotate_bearing(knot_c, rt.x, rt.y, bx, by, bz, X, Y, Z);
          project_jac(cam, X, Y, Z, M);
          vc[0] = h0 * M[0] + h1 * M[3];  // temp * dpm_ddrot (model.cpp:449)
          vc[1] = h0 * M[1];
          vc[2] = h0 * M[2] + h1 * M[5];
          row_times_A(knot_c, jt.x, jt.y, jt.z, vc, wc);
        }
        {
          const double4 rt = ldg256(RotTab + bp);
          const double4 jt = ldg256(JacTab + bp);
          double X, Y, Z, M[6];
          rotate_bearing(knot_p, rt.x, rt.y, bx, by, bz, X, Y, Z);
          project_jac(cam, X, Y, Z, M);
          vp[0] = -(g.x * M[0] + g.y * M[3]);  // -Gpm * dpm_ddrot (model.cpp:459)
          vp[1] = -(g.x * M[1]);
          vp[2] = -(g.x * M[2] + g.y * M[5]);
          row_times_A(knot_p, jt.x, jt.y, jt.z, vp, wp);
        }
        // IRLS weight (model.cpp:599-618), applied as sqrt(w) on the whole row and on e
        double sw = 1.0;
        if (COST == EMBA_COST_CAUCHY) sw = sqrt(1.0 / (1.0 + eta * e * e));
        if (COST == EMBA_COST_HUBER) { const double ab = fabs(e); sw = ab < eta ? 1.0 : sqrt(eta / ab); }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          row[k] = sw * (vc[k] - wc[k]);   // d/d(pose cp_c)   : [I - A | A] (so3_spline.h:254-269)
          row[3 + k] = sw * wc[k];         // d/d(pose cp_c+1)
          row[6 + k] = sw * (vp[k] - wp[k]);
          row[9 + k] = sw * wp[k];
        }
        row[12] = sw * e;
        d0 = sw * dpv.x;
        d1 = sw * dpv.y;
        // pose window of the pixel: one 8-byte look first, the integer atomics only when the window really grows.
        // Windows only grow, so a stale look costs at most a redundant atomic, never a missed one.
        const int2 w = win[a];
        if (it.cp_p < w.x) atomicMin(&win[a].x, it.cp_p);
        if (it.cp_c + 1 > w.y) atomicMax(&win[a].y, it.cp_c + 1);
      }
    }
    // the previous tile's TMA store must have drained the tile (and every warp finished reading it)
    if (store_pending && tid == 0) tma_store_wait_read();
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 6; c++) trow[c ^ swz] = make_double2(row[2 * c], row[2 * c + 1]);
    trow[6 ^ swz] = make_double2(row[12], d0);
    trow[7 ^ swz] =
        make_double2(d1, __longlong_as_double((long long)(meta_lo | ((unsigned long long)(uint32_t)m << 32))));
    fence_async_smem();
    __syncthreads();
    fetch(t0 + kAsmThreads);
    // Jacobian rows -> global
    const int nrec = min(kAsmThreads, it.count - t0);
    if (nrec == kAsmThreads) {
      if (tid == 0) tma_store_rows(&jmap, tile, it.start + t0);
      store_pending = true;
    } else {
      double2* out = reinterpret_cast<double2*>(jrec + ((size_t)it.start + t0) * kRecDoubles);
      const double2* t2 = reinterpret_cast<const double2*>(tile);
      for (int i = tid; i < nrec * 8; i += kAsmThreads) {
        const int r = i >> 3, c = i & 7;
        out[i] = t2[r * 8 + (c ^ (r & 7))];
      }
    }
    // rank-32 update of the three 8x8 blocks from this warp's own rows
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const int r = 32 * warp + 8 * (s >> 1) + (s & 1) + 2 * kk;
      const int rs = r & 7;
      const double f0 = tile[r * kRecDoubles + (((fq >> 1) ^ rs) << 1) + (fq & 1)];
      const double f1 = tile[r * kRecDoubles + ((((fq >> 1) + 4) ^ rs) << 1) + (fq & 1)];
      dmma884(c00a, c00b, f0, f0);
      dmma884(c01a, c01b, f0, f1);
      dmma884(c11a, c11b, f1, f1);
    }
  }
  // reduce the per-warp blocks in a fixed order (the tile doubles as scratch once the last store has read it)
  if (store_pending && tid == 0) tma_store_wait_read();
  __syncthreads();
  {
    double* red = tile + warp * 192;
    const int o = fq * 8 + 2 * kk;
    red[o] = c00a; red[o + 1] = c00b;
    red[64 + o] = c01a; red[64 + o + 1] = c01b;
    red[128 + o] = c11a; red[128 + o + 1] = c11b;
  }
  __syncthreads();
  if (tid < kAccN) {
    // tri13 order: row-major upper triangle of the 13 x 13 matrix
    int i = 0, k = tid;
    while (k >= 13 - i) { k -= 13 - i; i++; }
    const int jj = i + k;
    const int blk = i < 8 ? (jj < 8 ? 0 : 1) : 2;
    const int o = blk * 64 + (i & 7) * 8 + (jj & 7);
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < kAsmThreads / 32; w++) x += tile[w * 192 + o];
    acc_part[(size_t)blockIdx.x * kAccN + tid] = x;
  }
}

// TMA descriptor of the Jacobian-row buffer: [Mc][16] fp64, box 256 rows x 16, 128-byte swizzle. Rebuilt only
// when the buffer moves or the window's measurement count changes. cuTensorMapEncodeTiled comes from the
// driver through the runtime's entry-point query (no link-time libcuda dependency).
static int jrec_tensor_map(Handle* h) {
  if (h->jrec_tmap_ptr == h->d_jrec && h->jrec_tmap_rows == h->Mc) return EMBA_OK;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn ||
        qr != cudaDriverEntryPointSuccess) {
      h->err = "cuTensorMapEncodeTiled is not available from this driver";
      return EMBA_E_SUPPORT;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)kRecDoubles, (cuuint64_t)h->Mc};
  const cuuint64_t strides[1] = {(cuuint64_t)kRecDoubles * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)kRecDoubles, (cuuint32_t)kAsmThreads};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(h->jrec_tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                            h->d_jrec, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return EMBA_E_CUDA;
  }
  h->jrec_tmap_ptr = h->d_jrec;
  h->jrec_tmap_rows = h->Mc;
  return EMBA_OK;
}

// per-group sums of the item partials, fixed order
__global__ void k_group_sum(const double* __restrict__ acc_part, const int32_t* __restrict__ item0, int n_groups,
                            double* __restrict__ gsum) {
  const int g = blockIdx.x;
  const int k = threadIdx.x;
  if (g >= n_groups || k >= kAccN) return;
  double s = 0;
  for (int i = item0[g]; i < item0[g + 1]; i++) s += acc_part[(size_t)i * kAccN + k];
  gsum[(size_t)g * kAccN + k] = s;
}

// Dense A11 / b1 from the group sums (model.cpp:453-477): one CTA per control pose `a` owns rows 3a..3a+2 and
// gathers, in a fixed order, every group whose pose set {c, c+1, p, p+1} contains a.
__global__ void k_a11_gather(int n, int dmax, const int32_t* __restrict__ gid, const double* __restrict__ gsum,
                             double* __restrict__ A11, double* __restrict__ b1) {
  extern __shared__ double rowbuf[];  // [3][3n] + [3]
  const int a = blockIdx.x;
  const int tid = threadIdx.x;
  const int d3 = 3 * n;
  for (int i = tid; i < 3 * d3 + 3; i += blockDim.x) rowbuf[i] = 0.0;
  __syncthreads();
  auto visit = [&](int c, int d) {
    if (c < 0 || c > n - 2 || d < 0 || d > dmax || d > c) return;
    const int g = gid[(size_t)c * (dmax + 1) + d];
    if (g < 0) return;
    const int p = c - d;
    const int S[4] = {c, c + 1, p, p + 1};
    const double* Z = gsum + (size_t)g * kAccN;
    if (tid < 12) {
      const int r = tid < 9 ? tid / 3 : tid - 9;
      const int k = tid % 3;
      for (int s = 0; s < 4; s++) {
        if (S[s] != a) continue;
        const int i = 3 * s + r;
        if (tid < 9) {
          for (int t = 0; t < 4; t++) {
            const int j = 3 * t + k;
            const double z = i <= j ? Z[tri13(i, j)] : Z[tri13(j, i)];
            rowbuf[r * d3 + 3 * S[t] + k] += z;
          }
        } else {
          rowbuf[3 * d3 + r] += Z[tri13(i, 12)];
        }
      }
    }
  };
  for (int c = a - 1; c <= a; c++)
    for (int d = 0; d <= dmax; d++) visit(c, d);
  for (int p = a - 1; p <= a; p++) {
    if (p < 0) continue;
    for (int c = max(p, a + 1); c <= min(n - 2, p + dmax); c++) visit(c, c - p);
  }
  __syncthreads();
  for (int i = tid; i < 3 * d3; i += blockDim.x) A11[(size_t)(3 * a + i / d3) * d3 + (i % d3)] = rowbuf[i];
  if (tid < 3) b1[3 * a + tid] = rowbuf[3 * d3 + tid];
}

// ---------------------------------------------------------------------------------------------------
// pose windows -> strip lengths
__global__ void k_strip_len(const int2* __restrict__ win, int32_t* __restrict__ winlo, int32_t* __restrict__ winhi, int64_t Np,
                            int64_t* __restrict__ len) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  const int2 w = win[a];
  const int32_t lo = w.x, hi = w.y;
  winlo[a] = lo;
  winhi[a] = hi;
  len[a] = hi >= lo ? (int64_t)(hi - lo + 1) : 0;
}

// Map side: one warp per active pixel, rows in fixed (sorted) order. The 128-byte rows are gathered with cp.async
// (16 B per lane, 4 rows per instruction) into a 2-stage shared-memory ring (4 CTAs = 32 warps per SM), so the DRAM latency of the gather
// overlaps the reduction of the previous tiles. Lanes 0..23 own one A12 component (slot, row, col), lanes 24..28
// own A22 xx, xy, yy and b2 x, y. A12 components are accumulated per run of equal (cp_c, cp_p) -- rows are sorted
// by row id, i.e. grouped by control-pose pair -- and flushed into the pixel's strip in shared memory.
constexpr int kPixWarps = 8;
constexpr int kStripCap = 64;  // poses per shared-memory strip
constexpr int kPixTile = 16;   // rows per stage
constexpr int kPixStages = 2;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// one pixel: rows [seg0, seg1) of the sorted list. SMEM: the strip accumulates in shared memory (typed shared
// accesses) and is copied out at the end; otherwise (pose window longer than kStripCap) directly in global memory.
// gs: the strip's place in the local strip buffer; go: where the finished strip goes (== gs on one GPU; the owner
// rank's receive buffer, possibly over NVLink, in the peer-memory exchange -- stores only, never read back).
template <bool SMEM>
__device__ __forceinline__ void pix_one(int64_t seg0, int64_t seg1, int qlo, int len, double* __restrict__ gs, double* go,
                                        double* s_tile, double* s_strip, const uint32_t* __restrict__ sval,
                                        const double* __restrict__ jrec, int lane, int s, int r, int c, int ia,
                                        int ib, int group, double& acc_out, unsigned long long& mask0, unsigned long long& mask1) {
  double* sp = SMEM ? s_strip : gs;
  for (int i = lane; i < len * 6; i += 32) sp[i] = 0.0;
  const int lrow = lane >> 3, lchunk = lane & 7;  // cp.async role: 4 rows per instruction, 8 x 16 B per row
  const int ntiles = (int)((seg1 - seg0 + kPixTile - 1) / kPixTile);
  // row indices of the next tile to be issued are fetched one step ahead, so the index load and the row gather it
  // feeds are not two DRAM latencies in series
  auto load_idx = [&](int t) -> uint32_t {
    const int64_t base = seg0 + (int64_t)t * kPixTile;
    return (t < ntiles && base + lane < seg1 && lane < kPixTile) ? sval[base + lane] : 0u;
  };
  uint32_t idx_next = load_idx(0);
  auto issue = [&](int t) {
    const uint32_t mine = idx_next;
    idx_next = load_idx(t + 1);
    if (t < ntiles) {
      const int64_t base = seg0 + (int64_t)t * kPixTile;
      const int cnt = (int)min((int64_t)kPixTile, seg1 - base);
      double* dst = s_tile + (size_t)(t % kPixStages) * kPixTile * kRecDoubles;
#pragma unroll
      for (int k = 0; k < kPixTile / 4; k++) {
        const int row = 4 * k + lrow;
        const uint32_t m = __shfl_sync(0xffffffffu, mine, row);
        if (row < cnt) cp_async16(dst + row * kRecDoubles + 2 * lchunk, jrec + (size_t)m * kRecDoubles + 2 * lchunk);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int t = 0; t < kPixStages - 1; t++) issue(t);
  double acc = 0.0;
  uint32_t runkey = 0xFFFFFFFFu;
  bool have = false;
  const int lane_off = ((s & 1) - qlo) * 6 + r * 2 + c;
  auto flush = [&]() {
    const int pose = (int)(s >= 2 ? (runkey >> 16) : (runkey & 0xFFFFu));
    double* cell = sp + pose * 6 + lane_off;
    if (lane < 12) *cell += acc;
    __syncwarp();
    if (lane >= 12 && lane < 24) *cell += acc;
    __syncwarp();
    if (lane < 24) acc = 0.0;
  };
  for (int t = 0; t < ntiles; t++) {
    issue(t + kPixStages - 1);
    cp_async_wait<kPixStages - 1>();
    __syncwarp();
    const int cnt = (int)min((int64_t)kPixTile, seg1 - (seg0 + (int64_t)t * kPixTile));
    const double* tl = s_tile + (size_t)(t % kPixStages) * kPixTile * kRecDoubles;
    const double* pa = tl + ia;
    const double* pb = tl + ib;
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(tl + 15);
    int i = 0;
    for (; i + 4 <= cnt; i += 4) {
      const uint32_t k0 = pk[(i + 0) * 2 * kRecDoubles], k1 = pk[(i + 1) * 2 * kRecDoubles];
      const uint32_t k2 = pk[(i + 2) * 2 * kRecDoubles], k3 = pk[(i + 3) * 2 * kRecDoubles];
      const double a0 = pa[(i + 0) * kRecDoubles], b0 = pb[(i + 0) * kRecDoubles];
      const double a1 = pa[(i + 1) * kRecDoubles], b1 = pb[(i + 1) * kRecDoubles];
      const double a2 = pa[(i + 2) * kRecDoubles], b2v = pb[(i + 2) * kRecDoubles];
      const double a3 = pa[(i + 3) * kRecDoubles], b3 = pb[(i + 3) * kRecDoubles];
      if (have && k0 == runkey && k1 == runkey && k2 == runkey && k3 == runkey) {
        acc = fma(a0, b0, acc);
        acc = fma(a1, b1, acc);
        acc = fma(a2, b2v, acc);
        acc = fma(a3, b3, acc);
      } else {
        if (k0 != runkey || !have) { if (have) flush(); runkey = k0; have = true; }
        acc = fma(a0, b0, acc);
        if (k1 != runkey) { flush(); runkey = k1; }
        acc = fma(a1, b1, acc);
        if (k2 != runkey) { flush(); runkey = k2; }
        acc = fma(a2, b2v, acc);
        if (k3 != runkey) { flush(); runkey = k3; }
        acc = fma(a3, b3, acc);
      }
    }
    for (; i < cnt; i++) {
      const uint32_t k0 = pk[i * 2 * kRecDoubles];
      if (k0 != runkey || !have) { if (have) flush(); runkey = k0; have = true; }
      acc = fma(pa[i * kRecDoubles], pb[i * kRecDoubles], acc);
    }
    __syncwarp();
  }
  cp_async_wait<0>();
  if (have) flush();
  __syncwarp();
  if (SMEM) {
    for (int i = lane; i < len * 6; i += 32) go[i] = sp[i];
  } else if (go != gs) {
    for (int i = lane; i < len * 6; i += 32) go[i] = gs[i];
  }
  if (SMEM) {
    strip_mask_warp(sp, len, qlo, group, lane, mask0, mask1);  // occupancy of the finished strip (emba_internal.cuh)
  } else {
    // a long strip lives in global memory: re-reading it here would cost its whole size again inside the assembly
    // pass. The solve fills these masks in before its first Schur product (k_strip_mask, min_len = kStripCap + 1).
    mask0 = ~0ull;
    mask1 = ~0ull;
  }
  acc_out = acc;
}

__global__ void __launch_bounds__(kPixWarps * 32)
k_pix(int64_t a_begin, int64_t a_end, const int32_t* __restrict__ segoff, const int32_t* __restrict__ segend,
      const uint32_t* __restrict__ sval,
      const double* __restrict__ jrec, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
      const int64_t* __restrict__ stripoff, double* __restrict__ strip, const int32_t* __restrict__ apix,
      const double* __restrict__ Gx, const double* __restrict__ Gy, double alpha, double* __restrict__ A22,
      double* __restrict__ b2, int group, unsigned long long* __restrict__ gmask, int strip_cap,
      const int64_t* __restrict__ dst) {
  extern __shared__ __align__(16) unsigned char pix_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t smem_per_warp = (size_t)(kPixStages * kPixTile * kRecDoubles + strip_cap * 6) * 8;
  double* s_tile = reinterpret_cast<double*>(pix_smem + (size_t)warp * smem_per_warp);  // [stage][row][16]
  double* s_strip = s_tile + kPixStages * kPixTile * kRecDoubles;
  const int64_t nw = (int64_t)gridDim.x * kPixWarps;
  const int s = lane / 6, r = (lane % 6) >> 1, c = lane & 1;
  // record fields: 0..11 Jc|Jp, 12 e, 13..14 dp, 15 meta
  int ia = 15, ib = 15;
  if (lane < 24) { ia = 3 * s + r; ib = 13 + c; }
  else if (lane == 24) { ia = 13; ib = 13; }
  else if (lane == 25) { ia = 13; ib = 14; }
  else if (lane == 26) { ia = 14; ib = 14; }
  else if (lane == 27) { ia = 13; ib = 12; }
  else if (lane == 28) { ia = 14; ib = 12; }
  for (int64_t a = a_begin + (int64_t)blockIdx.x * kPixWarps + warp; a < a_end; a += nw) {
    const int64_t seg0 = segoff[a];
    const int64_t seg1 = segend[a];
    const int qlo = winlo[a];
    const int len = winhi[a] >= qlo ? winhi[a] - qlo + 1 : 0;  // empty window: no local rows (multi-GPU)
    double* gs = strip + stripoff[a] * 6;
    double* go = dst ? reinterpret_cast<double*>((uintptr_t)dst[a]) : gs;
    double acc = 0.0;
    unsigned long long mask0 = 0ull, mask1 = 0ull;
    if (len <= strip_cap) pix_one<true>(seg0, seg1, qlo, len, gs, go, s_tile, s_strip, sval, jrec, lane, s, r, c, ia, ib, group, acc, mask0, mask1);
    else pix_one<false>(seg0, seg1, qlo, len, gs, go, s_tile, s_strip, sval, jrec, lane, s, r, c, ia, ib, group, acc, mask0, mask1);
    if (lane == 0) { gmask[2 * a] = mask0; gmask[2 * a + 1] = mask1; }
    // applyL2Reg (model.cpp:689-719): A22 += alpha*I, b2 -= alpha * (Gx, Gy)[pixel]
    const int32_t pix = apix[a];
    if (lane == 24) A22[3 * a] = acc + alpha;
    if (lane == 25) A22[3 * a + 1] = acc;
    if (lane == 26) A22[3 * a + 2] = acc + alpha;
    if (lane == 27) b2[2 * a] = acc - alpha * Gx[pix];
    if (lane == 28) b2[2 * a + 1] = acc - alpha * Gy[pix];
    __syncwarp();
  }
  if (dst) __threadfence_system();  // remote sub-strips are performed before the kernel counts as finished
}

// occupancy masks of finished strips with at least min_len poses, one warp per pixel (the solve calls it for the
// strips k_pix kept in global memory, or for all of them after the fp64-atomic path, which has no k_pix)
__global__ void k_strip_mask(int64_t Np, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
                             const int64_t* __restrict__ stripoff, const double* __restrict__ strip, int group,
                             unsigned long long* __restrict__ gmask, int min_len) {
  const int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= Np) return;
  const int lo = winlo[a], hi = winhi[a];
  const int len = hi >= lo ? hi - lo + 1 : 0;
  if (len < min_len) return;
  unsigned long long m0, m1;
  strip_mask_warp(strip + stripoff[a] * 6, len, lo, group, lane, m0, m1);
  if (lane == 0) { gmask[2 * a] = m0; gmask[2 * a + 1] = m1; }
}

// peer-memory exchange behind the fp64-atomic map path: one warp per pixel copies its finished local strip to the
// owner's receive buffer
__global__ void k_strip_push(int64_t Np, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
                             const int64_t* __restrict__ stripoff, const double* __restrict__ strip,
                             const int64_t* __restrict__ dst) {
  const int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= Np) return;
  const int lo = winlo[a], hi = winhi[a];
  const int len = hi >= lo ? hi - lo + 1 : 0;
  const double* src = strip + stripoff[a] * 6;
  double* out = reinterpret_cast<double*>((uintptr_t)dst[a]);
  for (int i = lane; i < len * 6; i += 32) out[i] = src[i];
  __threadfence_system();
}

int fill_strip_masks(Handle* h) {
  if (h->mask_min_len < 0 || h->Np <= 0) return EMBA_OK;
  k_strip_mask<<<ceil_div64(h->Np * 32, 256), 256, 0, h->stream>>>(h->Np, h->sv_winlo, h->sv_winhi, h->sv_stripoff,
                                                                  h->sv_strip, h->pose_group, h->sv_gmask,
                                                                  h->mask_min_len);
  EMBA_LAUNCH_CHECK();
  h->mask_min_len = -1;
  return EMBA_OK;
}

// fp64-atomic map-block path (reported beside the deterministic one): one thread per Jacobian row in canonical
// order, 24 + 5 atomic adds into the pixel's strip / A22 / b2. Strips, A22 and b2 must be zeroed first.
__global__ void __launch_bounds__(256)
k_map_atomic(int64_t Mc, const int32_t* __restrict__ pix_in, const int32_t* __restrict__ amap,
             const double* __restrict__ jrec,
             const int32_t* __restrict__ winlo, const int64_t* __restrict__ stripoff, double* __restrict__ strip,
             double* __restrict__ A22, double* __restrict__ b2) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mc) return;
  const int32_t pix = pix_in[m];
  const int32_t a = pix >= 0 ? amap[pix] : -1;
  if (a < 0) return;
  const double4* r4 = reinterpret_cast<const double4*>(jrec + (size_t)m * kRecDoubles);
  const double4 q0 = ldg256(r4), q1 = ldg256(r4 + 1), q2 = ldg256(r4 + 2), q3 = ldg256(r4 + 3);
  const double J[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
  const double e = q3.x, d0 = q3.y, d1 = q3.z;
  const uint32_t key = (uint32_t)(unsigned long long)__double_as_longlong(q3.w);
  const int cpc = (int)(key & 0xFFFFu), cpp = (int)(key >> 16);
  double* sp = strip + (stripoff[a] - winlo[a]) * 6;
  const int pose[4] = {cpc, cpc + 1, cpp, cpp + 1};
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double* c = sp + pose[s] * 6 + r * 2;
      atomicAdd(c, J[3 * s + r] * d0);
      atomicAdd(c + 1, J[3 * s + r] * d1);
    }
  atomicAdd(&A22[3 * (size_t)a], d0 * d0);
  atomicAdd(&A22[3 * (size_t)a + 1], d0 * d1);
  atomicAdd(&A22[3 * (size_t)a + 2], d1 * d1);
  atomicAdd(&b2[2 * (size_t)a], d0 * e);
  atomicAdd(&b2[2 * (size_t)a + 1], d1 * e);
}

__global__ void k_l2_reg(int64_t Np, const int32_t* __restrict__ apix, const double* __restrict__ Gx,
                         const double* __restrict__ Gy, double alpha, double* __restrict__ A22,
                         double* __restrict__ b2) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  const int32_t pix = apix[a];
  A22[3 * a] += alpha;
  A22[3 * a + 2] += alpha;
  b2[2 * a] -= alpha * Gx[pix];
  b2[2 * a + 1] -= alpha * Gy[pix];
}

// dense copy of A12 for emba_get_normal_eq (parity only)
__global__ void k_a12_dense(int64_t Np, int n, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
                            const int64_t* __restrict__ stripoff, const double* __restrict__ strip,
                            double* __restrict__ out) {
  const int64_t a = blockIdx.x;
  if (a >= Np) return;
  const int lo = winlo[a], hi = winhi[a];
  if (hi < lo) return;
  const double* sp = strip + stripoff[a] * 6;
  for (int i = threadIdx.x; i < (hi - lo + 1) * 6; i += blockDim.x) {
    const int q = lo + i / 6, rr = (i % 6) >> 1, cc = i & 1;
    out[(size_t)(3 * q + rr) * (2 * Np) + 2 * a + cc] = sp[i];
  }
}
__global__ void k_a22_expand(int64_t Np, const double* __restrict__ A22, double* __restrict__ out) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  out[4 * a] = A22[3 * a];
  out[4 * a + 1] = A22[3 * a + 1];
  out[4 * a + 2] = A22[3 * a + 1];
  out[4 * a + 3] = A22[3 * a + 2];
}
__global__ void k_i32_to_i64(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

int form_normal_eq(Handle* h, int thres, int cost_type, double eta, double alpha) {
  StateSlot& s = h->st[h->cur];
  const int64_t P = h->P;
  const int T = 256;
  const int n = h->n;
  h->thres = thres;
  h->formed = false;
  h->pose_group = 16 * ((n + 1023) / 1024);  // at most 64 groups
  EMBA_CUDA(cudaEventRecord(h->ev[4], h->stream));
  const bool atomic_path = h->map_path == EMBA_MAP_ATOMIC;
  uint32_t* vs = h->d_sval;
  h->jrec_valid = false;
  // ---- 1. active set (model.cpp:324-379): mask + exclusive scan reproduces the ascending std::set order. The same
  // pass sizes the row segment of every active pixel from the evaluation's (rank-local) histogram.
  int32_t *d_flag = h->d_pflag, *d_aidx = h->d_paidx;
  int32_t* d_rowbase = h->d_segcnt;
  int64_t* d_len = h->d_len;
  void* scan_tmp2 = reinterpret_cast<char*>(h->d_scan_tmp) + scan_scratch_bytes(P + 2);
#define EMBA_TRYC(x) do { int _r = (x); if (_r != EMBA_OK) return _r; } while (0)
#define EMBA_CUDAC(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(_e); return EMBA_E_CUDA; } } while (0)
  int64_t* d_totals = reinterpret_cast<int64_t*>(h->d_scal + 32);
  EMBA_CUDAC(cudaMemsetAsync(d_totals, 0, 2 * sizeof(int64_t), h->stream));
  k_active_flags<<<ceil_div64(P, T), T, 0, h->stream>>>(s.hist, s.hist_loc, P, thres, d_flag, d_rowbase);
  h->launches++;
  EMBA_TRYC(scan_exclusive<int32_t>(h, h->stream, d_flag, d_aidx, P, h->d_scan_tmp));
  EMBA_TRYC(scan_exclusive<int32_t>(h, h->stream, d_rowbase, d_rowbase, P, scan_tmp2));
  // Np goes back to the host through pinned memory; the host waits for it only after the pose-side kernel (which
  // does not need it) has been queued, so the round trip hides behind that kernel
  k_active_fill<<<ceil_div64(P, T), T, 0, h->stream>>>(d_flag, d_aidx, s.hist_loc, d_rowbase, P, h->d_amap, h->d_apix,
                                                      h->d_win64, s.H3, h->d_segoff, h->d_segend, d_totals);
  h->launches++;
  EMBA_CUDAC(cudaMemcpyAsync(h->h_pin, d_totals, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDAC(cudaEventRecord(h->ev_host, h->stream));
  // ---- 2. pose side + Jacobian rows
  const int64_t Mc = h->Mc;
  const PanoCam cam = make_cam(h);
  EMBA_CUDAC(cudaEventRecord(h->ev_fork, h->stream));  // the side stream's kernels need everything up to here
  EMBA_CUDAC(cudaEventRecord(h->ev[5], h->stream));
  if (h->n_items > 0) {
    EMBA_TRYC(jrec_tensor_map(h));
    // (Capping this HBM-bound kernel at 2 CTAs per SM with a shared-memory pad, to leave a third of every SM to the
    // side stream's placement / segment sort, was measured on C4: form 14.3 ms against 13.3 ms without the pad. What
    // matters is the submission order -- this kernel first, the side stream's kernels behind it: they then fill the
    // SMs as its CTAs retire instead of delaying its start. EMBA_ASM_PAD keeps the experiment reachable.)
    static const int asm_pad = getenv("EMBA_ASM_PAD") ? atoi(getenv("EMBA_ASM_PAD")) : 0;
    if (asm_pad > 48 * 1024 - 33 * 1024) {
      EMBA_CUDAC(cudaFuncSetAttribute(k_asm_pose<EMBA_COST_QUADRATIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, asm_pad));
      EMBA_CUDAC(cudaFuncSetAttribute(k_asm_pose<EMBA_COST_CAUCHY>, cudaFuncAttributeMaxDynamicSharedMemorySize, asm_pad));
      EMBA_CUDAC(cudaFuncSetAttribute(k_asm_pose<EMBA_COST_HUBER>, cudaFuncAttributeMaxDynamicSharedMemorySize, asm_pad));
    }
    const CUtensorMap jmap = *reinterpret_cast<const CUtensorMap*>(h->jrec_tmap);
#define EMBA_ASM_LAUNCH(C)                                                                                         \
  k_asm_pose<C><<<h->n_items, kAsmThreads, asm_pad, h->stream>>>(jmap, h->d_items, h->d_rec, s.Ktab, s.RotTab,          \
                                                           s.JacTab, s.G2,                                        \
                                                           s.H3, s.dp, s.e, s.pix, cam, eta,                       \
                                                           h->d_jrec,                                              \
                                                           h->d_win64, h->d_acc_part)
    if (cost_type == EMBA_COST_QUADRATIC) EMBA_ASM_LAUNCH(EMBA_COST_QUADRATIC);
    else if (cost_type == EMBA_COST_CAUCHY) EMBA_ASM_LAUNCH(EMBA_COST_CAUCHY);
    else EMBA_ASM_LAUNCH(EMBA_COST_HUBER);
#undef EMBA_ASM_LAUNCH
    h->launches++;
    EMBA_CUDAC(cudaGetLastError());
  }
  EMBA_CUDAC(cudaEventRecord(h->ev[6], h->stream));  // end of the pose-side kernel
  // ---- 1b. side stream: rows -> pixel segments (k_place) and the per-segment ordering (k_seg_sort*). They need only
  // the evaluation and the segment offsets; they are submitted BEHIND the pose-side kernel (see above) and the main
  // stream joins them before the map-side kernel.
  if (!atomic_path && h->Mc > 0) {
    const int64_t Mc = h->Mc;
    // EMBA_SIDE_SERIAL=1 (read per call; bench.py's per-kernel roofline leg): the same kernels on the main stream, behind
    // the pose-side kernel, so that every kernel's event-timed duration is its standalone one. The pass takes the same
    // time either way: both sides are bound by the memory system, what the side stream overlaps it takes from the
    // pose-side kernel (C4: 7.0 ms beside the side stream, 5.2 ms alone).
    const bool side_serial = getenv("EMBA_SIDE_SERIAL") && atoi(getenv("EMBA_SIDE_SERIAL")) == 1;
    cudaStream_t sst = side_serial ? h->stream : h->stream2;
    EMBA_CUDAC(cudaStreamWaitEvent(sst, h->ev_fork, 0));
    EMBA_CUDAC(cudaEventRecord(h->ev_sort0, sst));
    EMBA_CUDAC(cudaMemsetAsync(h->d_longlist, 0, sizeof(int32_t), sst));
    const int place_v = getenv("EMBA_PLACE_V") ? atoi(getenv("EMBA_PLACE_V")) : 0;  // read per call (tools/place_variants.py)
    if (place_v == 1) k_place<4, true><<<ceil_div64(Mc, 256 * 4), 256, 0, sst>>>(Mc, s.pix, s.slot, d_rowbase, h->d_sval);
    else if (place_v == 2) k_place<8, false><<<ceil_div64(Mc, 256 * 8), 256, 0, sst>>>(Mc, s.pix, s.slot, d_rowbase, h->d_sval);
    else if (place_v == 3) k_place<8, true><<<ceil_div64(Mc, 256 * 8), 256, 0, sst>>>(Mc, s.pix, s.slot, d_rowbase, h->d_sval);
    else if (place_v == 4) k_place<2, true><<<ceil_div64(Mc, 256 * 2), 256, 0, sst>>>(Mc, s.pix, s.slot, d_rowbase, h->d_sval);
    else k_place<4, false><<<ceil_div64(Mc, 256 * 4), 256, 0, sst>>>(Mc, s.pix, s.slot, d_rowbase, h->d_sval);
    k_seg_sort<<<h->sm_count * 32, kSegWarps * 32, 0, sst>>>(d_totals, h->d_segoff, h->d_segend, h->d_sval, h->d_longlist);
    k_seg_sort_long<<<h->sm_count * 16, kSegLongWarps * 32, 0, sst>>>(h->d_longlist, h->d_segoff, h->d_segend, h->d_sval);
    const int huge_smem = 200 * 1024;
    EMBA_CUDAC(cudaFuncSetAttribute(k_seg_sort_huge, cudaFuncAttributeMaxDynamicSharedMemorySize, huge_smem));
    k_seg_sort_huge<<<h->sm_count, 512, huge_smem, sst>>>(h->d_longlist, h->d_segoff, h->d_segend, h->d_sval, huge_smem / 4);
    h->launches += 4;
    EMBA_CUDAC(cudaGetLastError());
    EMBA_CUDAC(cudaEventRecord(h->ev_sort1, sst));
    EMBA_CUDAC(cudaEventRecord(h->ev_join, sst));
  }
  EMBA_CUDAC(cudaEventSynchronize(h->ev_host));
  const int64_t Np = h->h_pin[0];
  h->Np = Np;
  h->Ma = h->h_pin[1];
  // ---- 3. map side: pose windows -> strip offsets. The strip total is read back while the A11 / b1 gather runs.
  EMBA_CUDAC(cudaMemsetAsync(d_len, 0, sizeof(int64_t) * (Np + 1), h->stream));
  if (Np) { k_strip_len<<<ceil_div64(Np, T), T, 0, h->stream>>>(h->d_win64, h->d_winlo, h->d_winhi, Np, d_len); h->launches++; }
  EMBA_TRYC(scan_exclusive<int64_t>(h, h->stream, d_len, h->d_stripoff, Np + 1, h->d_scan_tmp));
  EMBA_CUDAC(cudaMemcpyAsync(h->h_pin + 2, h->d_stripoff + Np, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDAC(cudaEventRecord(h->ev_host, h->stream));
  // the small A11 / b1 gather is latency-bound on a tiny grid: it goes to the side stream (behind the sort) and
  // runs beside the map-side kernel; the main stream joins it at the end
  EMBA_CUDAC(cudaEventRecord(h->ev_fork2, h->stream));
  EMBA_CUDAC(cudaStreamWaitEvent(h->stream2, h->ev_fork2, 0));
  EMBA_CUDAC(cudaMemsetAsync(h->d_A11, 0, sizeof(double) * 9 * n * n, h->stream2));
  EMBA_CUDAC(cudaMemsetAsync(h->d_b1, 0, sizeof(double) * 3 * n, h->stream2));
  if (h->n_groups > 0) {
    k_group_sum<<<h->n_groups, 96, 0, h->stream2>>>(h->d_acc_part, h->d_group_item0, h->n_groups, h->d_gsum);
    h->launches++;
    const size_t shm = sizeof(double) * (9 * (size_t)n + 3);
    if (shm > 48 * 1024) EMBA_CUDAC(cudaFuncSetAttribute(k_a11_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    k_a11_gather<<<n, 64, shm, h->stream2>>>(n, h->dmax, h->d_gid, h->d_gsum, h->d_A11, h->d_b1);
    h->launches++;
    EMBA_CUDAC(cudaGetLastError());
  }
  EMBA_CUDAC(cudaEventRecord(h->ev_join2, h->stream2));
  // multi-GPU: A11 / b1 stay partial here. The damped pose block A11 + lambda*diag(A11) is linear in A11, so every
  // rank subtracts its Schur contributions from its OWN partial and one all-reduce in the solve combines both
  // (saves a 72 n^2-byte all-reduce per assembly); emba_get_normal_eq combines them on demand.
  h->a11_partial = h->world > 1;
  // several GPUs: the pose windows are final now -- exchange them and size the strip transfers BEFORE the map-side
  // kernel, so that the transfers can overlap it (the host reads those numbers with the same synchronisation as
  // the strip total)
  const bool exchange = h->world > 1 && Np > 0 && !atomic_path;
  // The ring-pipelined variant (one map-side launch per owner's pixel range, each range leaving on a communication
  // stream while the next is reduced) is kept behind EMBA_XCHG_PIPELINE=1: measured on C4 it LOSES to one grouped
  // all-to-all after the map-side kernel (8 GPUs: pass 4.57 vs 3.16 ms; 4 GPUs: 5.77 vs 5.33 ms) -- seven small
  // send/recv launches with their rendezvous cost more than the transfer they hide, and the chunked map-side
  // launches lose 0.2 ms to their tails.
  static const bool pipe_env = getenv("EMBA_XCHG_PIPELINE") && atoi(getenv("EMBA_XCHG_PIPELINE")) == 1;
  bool pipelined = pipe_env && h->world <= 64;
  if (exchange) {
    EMBA_CUDAC(cudaEventRecord(h->ev_x[2], h->stream));
    EMBA_TRYC(comm_exchange_prepare(h));
    EMBA_CUDAC(cudaEventRecord(h->ev_host, h->stream));
    EMBA_CUDAC(cudaEventRecord(h->ev_x[3], h->stream));
  }
  EMBA_CUDAC(cudaEventSynchronize(h->ev_host));
  const int64_t tot = h->h_pin[2];
  h->strip_total = tot;
  h->peer_now = false;
  if (exchange) {
    EMBA_TRYC(comm_exchange_sizes(h));
    if (h->peer_now) { pipelined = false; EMBA_TRYC(comm_exchange_dst(h)); }
  }
  EMBA_TRYC(dev_reserve(h, &h->d_strip, &h->strip_cap, tot * 6 + tot * 3));  // +50 %: windows drift between iterations
  EMBA_CUDAC(cudaEventRecord(h->ev[8], h->stream));
  if (atomic_path) {
    EMBA_CUDAC(cudaEventRecord(h->ev[9], h->stream));
    EMBA_CUDAC(cudaMemsetAsync(h->d_strip, 0, sizeof(double) * 6 * (size_t)tot, h->stream));
    EMBA_CUDAC(cudaMemsetAsync(h->d_A22, 0, sizeof(double) * 3 * (size_t)Np, h->stream));
    EMBA_CUDAC(cudaMemsetAsync(h->d_b2, 0, sizeof(double) * 2 * (size_t)Np, h->stream));
    if (Mc > 0) {
      k_map_atomic<<<ceil_div64(Mc, 256), 256, 0, h->stream>>>(Mc, s.pix, h->d_amap, h->d_jrec, h->d_winlo,
                                                               h->d_stripoff, h->d_strip, h->d_A22, h->d_b2);
      h->launches++;
      EMBA_CUDAC(cudaGetLastError());
    }
    if (Np > 0) {
      StateSlot& sc = h->st[h->cur];
      k_l2_reg<<<ceil_div64(Np, 256), 256, 0, h->stream>>>(Np, h->d_apix, sc.Gx, sc.Gy, h->rank == 0 ? alpha : 0.0,
                                                          h->d_A22, h->d_b2);
      h->launches++;
    }
  }
  if (!atomic_path && Mc > 0) EMBA_CUDAC(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  if (!atomic_path) EMBA_CUDAC(cudaEventRecord(h->ev[9], h->stream));
  int strip_cap_used = kStripCap;
  if (Np > 0 && !atomic_path) {
    // shared-memory strip capacity per warp: 64 poses (4 CTAs/SM). Longer windows accumulate in global memory;
    // larger capacities were measured slower (C3, mean window 76 poses: 3.3 ms at 64, 3.7 / 4.4 / 3.9 ms at
    // 96 / 128 / 160) because of the occupancy they cost. EMBA_PIX_CAP overrides for experiments.
    static const int cap_env = getenv("EMBA_PIX_CAP") ? atoi(getenv("EMBA_PIX_CAP")) : 0;
    const int strip_cap = cap_env > 0 ? cap_env : kStripCap;
    strip_cap_used = strip_cap;
    const int pix_smem = kPixWarps * (kPixStages * kPixTile * kRecDoubles + strip_cap * 6) * 8;
    EMBA_CUDAC(cudaFuncSetAttribute(k_pix, cudaFuncAttributeMaxDynamicSharedMemorySize, pix_smem));
    const int ctas_per_sm = std::max(1, std::min(4, (227 * 1024) / (pix_smem + 1024)));
    auto launch_pix = [&](int64_t a_begin, int64_t a_end) -> int {
      if (a_end <= a_begin) return EMBA_OK;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a_end - a_begin + kPixWarps - 1) / kPixWarps, (int64_t)h->sm_count * 4 * ctas_per_sm));
      k_pix<<<grid, kPixWarps * 32, pix_smem, h->stream>>>(a_begin, a_end, h->d_segoff, h->d_segend, vs, h->d_jrec, h->d_winlo, h->d_winhi,
                                                    h->d_stripoff, h->d_strip, h->d_apix, s.Gx, s.Gy,
                                                    h->rank == 0 ? alpha : 0.0, h->d_A22, h->d_b2, h->pose_group, h->d_gmask, strip_cap,
                                                    h->peer_now ? h->d_dst : nullptr);
      h->launches++;
      EMBA_CUDAC(cudaGetLastError());
      return EMBA_OK;
    };
    if (!exchange || !pipelined) {
      EMBA_TRYC(launch_pix(0, Np));
    } else {
      // one launch per owner's pixel range, in ring order: the strips of rank r+1's pixels first -- they leave on the
      // communication stream while the next range is being reduced -- my own range last
      const int W = h->world, r = h->rank;
      for (int st = 1; st <= W; st++) {
        const int q = (r + st) % W;
        EMBA_TRYC(launch_pix(Np * q / W, Np * (q + 1) / W));
        if (st < W) {
          EMBA_CUDAC(cudaEventRecord(h->ev_chunk[st], h->stream));
          EMBA_CUDAC(cudaStreamWaitEvent(h->stream3, h->ev_chunk[st], 0));
          EMBA_TRYC(comm_exchange_step(h, st, h->stream3));
        }
      }
      EMBA_CUDAC(cudaEventRecord(h->ev_comm, h->stream3));
    }
  }
  EMBA_CUDAC(cudaEventRecord(h->ev[10], h->stream));
  EMBA_CUDAC(cudaStreamWaitEvent(h->stream, h->ev_join2, 0));
  h->sv_winlo = h->d_winlo; h->sv_winhi = h->d_winhi; h->sv_stripoff = h->d_stripoff; h->sv_strip = h->d_strip;
  h->sv_gmask = h->d_gmask;
  h->mask_min_len = atomic_path ? 0 : strip_cap_used + 1;  // strips whose masks the solve still has to fill in (-1: none)
  h->sv_strip_total = tot;
  if (h->world > 1 && Np > 0) {
    // A22 / b2 are small: all-reduce. A12: every rank's strips cover (almost) disjoint pose ranges, so they are not
    // summed everywhere; each rank becomes the owner of a contiguous range of pixels and receives only the
    // sub-strips of those pixels (1/world of the volume), see comm.cu
    if (!exchange) {  // fp64-atomic map path: nothing was overlapped, run the exchange back to back
      EMBA_TRYC(comm_exchange_prepare(h));
      EMBA_CUDAC(cudaStreamSynchronize(h->stream));
      EMBA_TRYC(comm_exchange_sizes(h));
      if (h->peer_now) {  // peer-memory exchange: the finished local strips are pushed to their owners by a copy kernel
        EMBA_TRYC(comm_exchange_dst(h));
        k_strip_push<<<ceil_div64(Np * 32, 256), 256, 0, h->stream>>>(Np, h->d_winlo, h->d_winhi, h->d_stripoff, h->d_strip, h->d_dst);
        h->launches++;
        EMBA_CUDAC(cudaGetLastError());
      }
    }
    if (!exchange || !pipelined) {
      EMBA_TRYC(comm_exchange_all(h, h->stream, true));
    } else {
      EMBA_CUDAC(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
    }
    EMBA_CUDAC(cudaEventRecord(h->ev_x[4], h->stream));
    EMBA_TRYC(comm_exchange_finish(h, !exchange || !pipelined));
    EMBA_CUDAC(cudaEventRecord(h->ev_x[5], h->stream));
  }
  EMBA_CUDAC(cudaEventRecord(h->ev[7], h->stream));
  EMBA_CUDAC(cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[4], h->ev[7]); h->t_ms[2] = ms;
  cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]); h->t_ms[3] = ms;
  cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]); h->t_ms[4] = ms;
  cudaEventElapsedTime(&ms, h->ev[9], h->ev[10]); h->t_ms[6] = ms;
  h->t_ms[7] = 0.0;
  if (!atomic_path && Mc > 0) { cudaEventElapsedTime(&ms, h->ev_sort0, h->ev_sort1); h->t_ms[7] = ms; }
  if (exchange) {
    cudaEventElapsedTime(&ms, h->ev_x[2], h->ev_x[3]); h->t_comm_ms[1] = ms;
    cudaEventElapsedTime(&ms, h->ev[9], h->ev_x[4]); h->t_comm_ms[2] = ms;
    cudaEventElapsedTime(&ms, h->ev_x[4], h->ev_x[5]); h->t_comm_ms[3] = ms;
  }
#undef EMBA_TRYC
#undef EMBA_CUDAC
  h->formed = true;
  h->jrec_valid = true;
  h->solved = false;
  return EMBA_OK;
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_form_normal_eq(emba_handle_t hh, int32_t thres, int32_t cost_type, double eta, double alpha,
                        int64_t* Np_out) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->st[h->cur].evaluated) { h->err = "emba_form_normal_eq: evaluate the current state first"; return EMBA_E_ARG; }
  if (cost_type < 0 || cost_type > 2) { h->err = "bad cost type"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_TRY(form_normal_eq(h, thres, cost_type, eta, alpha));
  if (Np_out) *Np_out = h->Np;
  return EMBA_OK;
}

int emba_set_map_path(emba_handle_t hh, int32_t mode) {
  Handle* h = (Handle*)hh;
  if (!h || (mode != EMBA_MAP_SORTED && mode != EMBA_MAP_ATOMIC)) return EMBA_E_ARG;
  h->map_path = mode;
  h->formed = h->solved = false;
  return EMBA_OK;
}

int emba_apply_l2_reg(emba_handle_t hh, double alpha) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->formed) { h->err = "emba_apply_l2_reg: no normal equations formed"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  if (h->Np > 0) {
    StateSlot& s = h->st[h->cur];
    k_l2_reg<<<ceil_div64(h->Np, 256), 256, 0, h->stream>>>(h->Np, h->d_apix, s.Gx, s.Gy, alpha, h->d_A22, h->d_b2);
    EMBA_LAUNCH_CHECK();
  }
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  h->solved = false;
  return EMBA_OK;
}

int emba_get_normal_eq(emba_handle_t hh, double* A11, double* b1, double* A22, double* b2, int64_t* active,
                       double* A12_dense) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->formed) { h->err = "emba_get_normal_eq: no normal equations formed"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  const int n = h->n;
  const int64_t Np = h->Np;
  if (h->a11_partial) {  // parity / adapter download with several GPUs: combine the partial pose blocks now
    EMBA_TRY(comm_allreduce(h, h->d_A11, (int64_t)9 * n * n, 1));
    EMBA_TRY(comm_allreduce(h, h->d_b1, (int64_t)3 * n, 1));
    h->a11_partial = false;
  }
  if (A11) EMBA_CUDA(cudaMemcpyAsync(A11, h->d_A11, sizeof(double) * 9 * n * n, cudaMemcpyDeviceToHost, h->stream));
  if (b1) EMBA_CUDA(cudaMemcpyAsync(b1, h->d_b1, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
  if (b2 && Np) EMBA_CUDA(cudaMemcpyAsync(b2, h->d_b2, sizeof(double) * 2 * Np, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  if (A22 && Np) {
    EMBA_TRY(dev_reserve(h, &h->d_cg, &h->cg_cap, 4 * Np));
    double* tmp = h->d_cg;
    k_a22_expand<<<ceil_div64(Np, 256), 256, 0, h->stream>>>(Np, h->d_A22, tmp);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(A22, tmp, sizeof(double) * 4 * Np, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = "emba_get_normal_eq: A22 copy failed"; return EMBA_E_CUDA; }
  }
  if (active && Np) {
    int64_t* tmp = nullptr;
    EMBA_TRY(dev_alloc(h, &tmp, Np));
    k_i32_to_i64<<<ceil_div64(Np, 256), 256, 0, h->stream>>>(h->d_apix, Np, tmp);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(active, tmp, sizeof(int64_t) * Np, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { h->err = "emba_get_normal_eq: active copy failed"; return EMBA_E_CUDA; }
  }
  if (A12_dense && Np) {
    double* tmp = nullptr;
    const int64_t cnt = (int64_t)3 * n * 2 * Np;
    EMBA_TRY(dev_alloc(h, &tmp, cnt));
    cudaMemsetAsync(tmp, 0, sizeof(double) * cnt, h->stream);
    k_a12_dense<<<(unsigned)Np, 64, 0, h->stream>>>(Np, n, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip, tmp);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(A12_dense, tmp, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { h->err = "emba_get_normal_eq: A12 copy failed"; return EMBA_E_CUDA; }
  }
  return EMBA_OK;
}

int emba_get_counters(emba_handle_t hh, int64_t* out) {
  Handle* h = (Handle*)hh;
  if (!h || !out) return EMBA_E_ARG;
  int32_t nlong = 0;
  if (h->formed && h->map_path == EMBA_MAP_SORTED && h->Mc > 0 && h->Np > 0) {
    cudaSetDevice(h->device);
    cudaMemcpy(&nlong, h->d_longlist, sizeof(int32_t), cudaMemcpyDeviceToHost);
  }
  out[0] = h->Mc; out[1] = h->Mc_total; out[2] = h->formed ? h->Np : 0; out[3] = h->formed ? h->Ma : 0;
  out[4] = h->formed ? h->strip_total * 6 : 0; out[5] = h->formed ? h->sv_strip_total * 6 : 0;
  out[6] = h->n_items; out[7] = nlong;
  return EMBA_OK;
}

int emba_a12_entries(emba_handle_t hh, int64_t* out) {
  Handle* h = (Handle*)hh;
  if (!h || !out) return EMBA_E_ARG;
  *out = h->formed ? h->sv_strip_total * 6 : 0;
  return EMBA_OK;
}

}  // extern "C"
