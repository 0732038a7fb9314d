// LEGM::solveNormalEq (Schur complement onto the control poses, reference src/emba/model.cpp:721-792) and
// LEGM::solveNormalEqCG (Jacobi-preconditioned CG on the full system, model.cpp:794-840; iteration as in
// Eigen/src/IterativeLinearSolvers/ConjugateGradient.h:26-96), plus the state update
// (Model::updateTraj model.cpp:22-53, LEGM::updateMap model.cpp:863-903).
//
// A12 lives on the device as per-pixel strips (3x2 block per control pose inside the pixel's pose window), so
//   S   = A11m - sum_a U_a C_a U_a^T,   C_a = (A22_a + lambda diag A22_a)^-1
//   rhs = b1   - sum_a U_a C_a b2_a
// is a block-sparse SYRK: tiles of S are accumulated in registers over the pixels whose window meets the tile.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include "emba_internal.cuh"
namespace cg = cooperative_groups;

namespace emba {

int comm_allreduce(Handle* h, void* buf, int64_t count, int dtype);
int fill_strip_masks(Handle* h);

// C_a = inverse of the damped 2x2 block (model.cpp:743-759)
__global__ void k_a22_inv(int64_t Np, const double* __restrict__ A22, double lambda, double* __restrict__ C) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= Np) return;
  const double xx = A22[3 * a], xy = A22[3 * a + 1], yy = A22[3 * a + 2];
  const double mxx = xx + lambda * xx, myy = yy + lambda * yy;
  const double det = mxx * myy - xy * xy;
  const double inv = 1.0 / det;
  C[3 * a] = myy * inv;
  C[3 * a + 1] = -xy * inv;
  C[3 * a + 2] = mxx * inv;
}

// ---------------------------------------------------------------------------------------------------
// Schur tiles. Extended index space [0, d]: rows 0..d-1 are pose unknowns (row = 3*(pose - fix) + r), row d is
// the right-hand-side "row" (value b2_a per pixel on the J side). Tile = 48 rows (16 poses). One CTA accumulates
// one (I <= J) tile pair over one chunk of pixels: sum_a W_I,a C_a W_J,a^T with W the pixel's strip rows (x2).
//   * pixels whose pose window misses either tile are compacted away per block of 128 (ballot order);
//   * the strip rows of 8 listed pixels per step go global -> shared with 16-byte cp.async (zero-filled outside
//     the pixel's window), double-buffered, so no registers or store instructions are spent on staging;
//   * the 48 x 48 x 16 products run on the fp64 tensor cores (m8n8k4): each of the 4 warps owns a 3 x 3 group of
//     8x8 blocks; the 2x2 factor C_a is applied to the I-side fragment in registers (one shuffle with the lane
//     holding the pixel's other column + 2 FMAs), so neither side needs a transformed copy in memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kST = 48;
constexpr int kSPix = 8;   // pixels per step (K = 16)
constexpr int kSchurThreads = 128;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
// 16-byte async copy; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_s() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_s() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int MINB>
__global__ void __launch_bounds__(kSchurThreads, MINB)
k_schur_tiles(int64_t own0, int64_t own1, int d, int fix, int nt, int Z, const int32_t* __restrict__ winlo,
              const int32_t* __restrict__ winhi, const int64_t* __restrict__ stripoff,
              const double* __restrict__ strip, const double* __restrict__ C, const double* __restrict__ b2,
              const unsigned long long* __restrict__ gmask, int group, double* __restrict__ Spart, int order) {
  // CTA order: `order` 0 = pixel chunk fastest (heaviest tile pairs first over the whole grid), 1 = tile pair fastest:
  // the CTAs in flight together then work on the SAME few pixel chunks, whose strips every tile pair re-reads -- with
  // chunks small enough they stay in the 126 MB L2 instead of being streamed from HBM once per pair. Within a chunk
  // the pairs still go diagonal by diagonal (pairs on and near the diagonal meet the most pose windows).
  const int npairs_ = nt * (nt + 1) / 2;
  const int pair_idx = order ? (int)(blockIdx.x % npairs_) : (int)blockIdx.y;
  const int z = order ? (int)(blockIdx.x / npairs_) : (int)blockIdx.x;
  int I = pair_idx, dgl = 0;
  while (I >= nt - dgl) { I -= nt - dgl; dgl++; }
  const int J = I + dgl;
  int pair_lin = dgl;  // index of (I, J) in the I-major numbering k_schur_finish uses
  for (int k = 0; k < I; k++) pair_lin += nt - k;
  // pixel chunks of the range this rank OWNS (all active pixels with one GPU): chunking the whole pixel range left
  // (world - 1) / world of the CTAs with nothing but empty windows and the rest with single-GPU-sized chunks
  const int64_t a0 = own0 + (own1 - own0) * z / Z, a1 = own0 + (own1 - own0) * (z + 1) / Z;
  const int rowI0 = I * kST, rowJ0 = J * kST;
  const bool Jhas_rhs = (d >= rowJ0 && d < rowJ0 + kST);
  // pose ranges covered by the tiles (poses are >= fix)
  const int pI0 = rowI0 / 3 + fix, pI1 = min(d - 1, rowI0 + kST - 1) / 3 + fix;
  const int pJ0 = rowJ0 / 3 + fix, pJ1 = min(d - 1, rowJ0 + kST - 1) / 3 + fix;
  // pose groups the tiles overlap (occupancy masks of the strips): a pixel with no entry in one of the tiles adds
  // nothing to this pair even when its window spans it
  // (masks are stored per gauge choice, so that bit k <-> poses [group*k + fix, group*(k+1) + fix): with group == 16
  // a tile is exactly one bit)
  auto bits = [&](int p0, int p1) -> unsigned long long {
    const int g0 = min(63, (p0 - fix) / group), g1 = min(63, (p1 - fix) / group);
    return (g1 >= 63 ? ~0ull : ((1ull << (g1 + 1)) - 1ull)) & ~((1ull << g0) - 1ull);
  };
  const unsigned long long bitsI = bits(pI0, pI1), bitsJ = bits(pJ0, pJ1);
  // [stage][side][pixel][row] as (column 0, column 1) pairs: k = 2*pixel + column
  __shared__ __align__(16) double2 V[2][2][kSPix][kST];
  // per listed pixel: first row of its window in the reduced system, number of rows, strip base, C, b2
  __shared__ int32_t m_row0[kSchurThreads], m_rows[kSchurThreads];
  __shared__ int64_t m_base[kSchurThreads];
  __shared__ double m_c[kSchurThreads][3];
  __shared__ double2 m_b[kSchurThreads];
  __shared__ int32_t wcount[kSchurThreads / 32];
  __shared__ int32_t nlist;
  const int tid = threadIdx.x;
  const int fg = (tid & 31) >> 2, fk = tid & 3;                // fragment row / k index of this lane
  const int rb0 = 3 * (tid >> 6), cb0 = 3 * ((tid >> 5) & 1);  // first 8-row / 8-column block of this warp
  double acc[3][3][2];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) acc[r][c][0] = acc[r][c][1] = 0.0;
  // staging roles: a step is 8 pixels x 2 sides x 48 rows of 16 bytes. Thread t < 96 owns ONE (side, row) and copies it
  // for the 8 pixels of the step: its row number, tile side and the right-hand-side test are fixed for the whole
  // kernel and the per-pixel list entries are warp-uniform (broadcast) shared loads. (With the items dealt out as
  // t + 128 q the index arithmetic of the copies was ~60 % of the kernel's instructions, 11 per DMMA.)
  const bool stager = tid < 2 * kST;
  const int s_side = tid / kST, s_rr = tid % kST;
  const int s_grow = (s_side ? rowJ0 : rowI0) + s_rr;
  const bool s_rhs = s_side && s_grow == d;   // the right-hand-side "row" of the J tile carries b2_a (plain store)
  const bool s_in = s_grow < d;
  const double2* strip2 = reinterpret_cast<const double2*>(strip);

  for (int64_t base = a0; base < a1; base += kSchurThreads) {
    // compact the pixels of this block whose window meets both tiles (ballot order -> deterministic)
    __syncthreads();
    const int64_t a = base + tid;
    bool ok = false;
    int lo = 0, hi = -1;
    if (a < a1) {
      lo = winlo[a]; hi = winhi[a];
      const unsigned long long occ = gmask[2 * a + (fix ? 1 : 0)];
      const bool mI = (rowI0 < d) && hi >= pI0 && lo <= pI1 && (occ & bitsI);
      const bool mJ = ((rowJ0 < d) && hi >= pJ0 && lo <= pJ1 && (occ & bitsJ)) || Jhas_rhs;
      ok = mI && mJ && hi >= lo;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((tid & 31) == 0) wcount[tid >> 5] = __popc(bal);
    // entries past the list must multiply the zero-filled rows by finite numbers
    m_c[tid][0] = 0.0; m_c[tid][1] = 0.0; m_c[tid][2] = 0.0;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < (tid >> 5); w++) off += wcount[w];
    if (ok) {
      const int l = off + __popc(bal & ((1u << (tid & 31)) - 1));
      m_row0[l] = 3 * (lo - fix);
      m_rows[l] = 3 * (hi - lo + 1);
      m_base[l] = stripoff[a] * 3;  // in double2 units: 3 rows per pose
      m_c[l][0] = C[3 * a]; m_c[l][1] = C[3 * a + 1]; m_c[l][2] = C[3 * a + 2];
      m_b[l] = make_double2(b2[2 * a], b2[2 * a + 1]);
    }
    if (tid == kSchurThreads - 1) nlist = off + __popc(bal);
    __syncthreads();
    const int nl = nlist;
    auto issue = [&](int l0, int st) {
      if (l0 < nl && stager) {
#pragma unroll
        for (int pl = 0; pl < kSPix; pl++) {
          const int l = l0 + pl;
          if (s_rhs) {
            V[st][1][pl][s_rr] = l < nl ? m_b[l] : make_double2(0.0, 0.0);
            continue;
          }
          const double2* src = strip2;
          int bytes = 0;
          if (l < nl && s_in) {
            const int rel = s_grow - m_row0[l];
            if (rel >= 0 && rel < m_rows[l]) { src = strip2 + m_base[l] + rel; bytes = 16; }
          }
          cp_async16_zfill(&V[st][s_side][pl][s_rr], src, bytes);
        }
      }
      cp_async_commit_s();
    };
    issue(0, 0);
    int st = 0;
    for (int l0 = 0; l0 < nl; l0 += kSPix, st ^= 1) {
      issue(l0 + kSPix, st ^ 1);
      cp_async_wait_s<1>();
      __syncthreads();
#pragma unroll
      for (int k0 = 0; k0 < 2 * kSPix; k0 += 4) {
        const int pl = (k0 >> 1) + (fk >> 1);
        const double* vi = reinterpret_cast<const double*>(&V[st][0][pl][0]) + (fk & 1);
        const double* vj = reinterpret_cast<const double*>(&V[st][1][pl][0]) + (fk & 1);
        const double cs = m_c[l0 + pl][2 * (fk & 1)], cx = m_c[l0 + pl][1];
        double av[3], bv[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const double raw = vi[2 * (8 * (rb0 + r) + fg)];
          const double other = __shfl_xor_sync(0xffffffffu, raw, 1);
          av[r] = raw * cs + other * cx;  // (W_I C_a)[row][column fk & 1]
        }
#pragma unroll
        for (int c = 0; c < 3; c++) bv[c] = vj[2 * (8 * (cb0 + c) + fg)];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 3; c++) dmma884(acc[r][c][0], acc[r][c][1], av[r], bv[c]);
      }
      __syncthreads();
    }
    cp_async_wait_s<0>();
  }
  double* out = Spart + ((size_t)z * npairs_ + pair_lin) * (kST * kST);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double* o = out + (8 * (rb0 + r) + fg) * kST + 8 * (cb0 + c) + 2 * fk;
      o[0] = acc[r][c][0];
      o[1] = acc[r][c][1];
    }
}

// S = A11m - sum_z partial tiles (fixed order); rhs = b1 - (...). A11m = A11 + lambda*diag(A11) (model.cpp:728-730).
__global__ void k_schur_finish(int d, int fix, int n, int nt, int npairs, int Z, const double* __restrict__ Spart,
                               const double* __restrict__ A11, const double* __restrict__ b1, double lambda,
                               double* __restrict__ S, double* __restrict__ rhs, double* __restrict__ T, int mode) {
  // mode 0: S = A11m - sum, rhs = b1 - sum (single GPU); mode 1: T = sum only; mode 2: S = A11m - T, rhs = b1 - T
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)d * (d + 1);
  if (idx >= total) return;
  const int i = (int)(idx / (d + 1)), j = (int)(idx % (d + 1));  // j == d -> rhs
  int ii = i, jj = j;
  if (jj != d && ii > jj) { const int t = ii; ii = jj; jj = t; }
  const int I = ii / kST, J = jj / kST;
  // pair index of (I, J), I <= J
  int pair = 0;
  for (int k = 0; k < I; k++) pair += nt - k;
  pair += J - I;
  const int ri = ii % kST, rj = jj % kST;
  double s = 0.0;
  if (mode == 2) {
    s = T[idx];
  } else {
    for (int z = 0; z < Z; z++) s += Spart[((size_t)z * npairs + pair) * (kST * kST) + ri * kST + rj];
    if (mode == 1) { T[idx] = s; return; }
  }
  // S is stored bordered, (d+1) x (d+1): row d holds the right-hand side, so the factorisation below carries the
  // forward substitution along (the last row of L becomes D^-1 L^-1 rhs)
  const int d3 = 3 * n;
  const size_t ld = (size_t)d + 1;
  if (j == d) {
    const double v = b1[3 * fix + i] - s;
    rhs[i] = v;
    S[(size_t)d * ld + i] = v;
    S[(size_t)i * ld + d] = 0.0;
    if (i == 0) S[(size_t)d * ld + d] = 1.0;  // placeholder pivot, never used
  } else {
    double a = A11[(size_t)(3 * fix + i) * d3 + 3 * fix + j];
    if (i == j) a += lambda * a;
    S[(size_t)i * ld + j] = a - s;
  }
}

// Blocked right-looking LDL^T (no pivoting) of the dense symmetric Schur complement, lower triangle in place.
// The reference uses Eigen's LDLT (model.cpp:789); S is the Schur complement of a damped SPD system.
// Per 32-column panel two launches:
//   k_ldlt_panel : every CTA factors the 32x32 diagonal block in shared memory (redundantly -- it is tiny), then one
//                  thread per row of its 64-row slab solves the panel row: y = L_row D (scratch W), L_row in place;
//   k_ldlt_update: trailing update A_ij -= sum_m y_im L_jm over 64x64 lower-triangular tiles, 4x4 register tiles.
// Then k_ldlt_subst (one CTA): blocked forward / diagonal / backward substitution.
constexpr int kNB = 32;

constexpr int kLdltSmemDoubles = 2 * kNB * (kNB + 1) + 64 * (kNB + 1);  // == 2 * 64 * (kNB + 1)

__device__ __forceinline__ void ldlt_panel_body(int d, int k0, int nb, double* __restrict__ S, double* __restrict__ W,
                                                int32_t* __restrict__ flags, int slab, int nslab_stride, int nslabs,
                                                double* __restrict__ sm, int ncheck) {
  double (*Dk)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(sm);
  double (*Li)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(sm + kNB * (kNB + 1));
  double (*As)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(sm + 2 * kNB * (kNB + 1));
  const int tid = threadIdx.x;
  __syncthreads();
  // diagonal block (padded with the identity past nb) and, beside it, the first slab of rows below the block:
  // the slab does not depend on the factorisation, so its load latency hides behind the serial part
  for (int e = tid; e < kNB * kNB; e += 256) {
    const int i = e >> 5, j = e & 31;
    Dk[i][j] = (i < nb && j < nb) ? (j <= i ? S[(size_t)(k0 + i) * d + k0 + j] : 0.0) : (i == j ? 1.0 : 0.0);
  }
  if (slab < nslabs) {
    const int r0 = k0 + nb + slab * 64;
    for (int e = tid; e < 64 * kNB; e += 256) {
      const int rr = e >> 5, j = e & 31;
      As[rr][j] = (r0 + rr < d && j < nb) ? S[(size_t)(r0 + rr) * d + k0 + j] : 0.0;
    }
  }
  __syncthreads();
  // The 32x32 block is factored by ONE warp with its rows in registers (lane i owns row i; loops fully unrolled,
  // static register indices): no block barriers on this serial critical path. Column j travels through a
  // 32-entry shared array read back as broadcasts. Entries above the diagonal accumulate unused junk.
  if (tid < 32) {
    const int lane = tid;
    double* ycol = &Li[0][0];
    double a[kNB];
#pragma unroll
    for (int m = 0; m < kNB; m++) a[m] = Dk[lane][m];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kNB; j++) {
      __syncwarp();
      ycol[lane] = a[j];  // y_ij = a_ij (lanes i >= j)
      __syncwarp();
      const double dj = ycol[j];
      // an exactly zero pivot (a control pose no active measurement touches: its whole row and column are zero) is
      // treated like Eigen's LDLT::solve treats it -- that component of the solution is 0 (model.cpp:789)
      bad = bad || (k0 + j < ncheck && !isfinite(dj));
      const double lij = dj != 0.0 ? a[j] / dj : 0.0;  // l_ij
#pragma unroll
      for (int m = j + 1; m < kNB; m++) a[m] -= lij * ycol[m];  // a_im -= l_ij * y_mj (used for j < m <= i)
      if (lane > j) a[j] = lij;
    }
    if (lane == 0 && slab == 0 && bad) atomicOr(flags, 8);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < kNB; m++)
      if (m <= lane) Dk[lane][m] = a[m];
    __syncwarp();
    // explicit inverse of the unit lower-triangular block: lane j builds column j by forward substitution,
    // x_i = [i == j] - sum_{m < i} l_im x_m (entries above the diagonal stay exactly zero); l_im is a broadcast
    double x[kNB];
#pragma unroll
    for (int i = 0; i < kNB; i++) {
      double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m + 1 < i; m += 2) {
        s0 -= Dk[i][m] * x[m];
        s1 -= Dk[i][m + 1] * x[m + 1];
      }
      if (i & 1) s0 -= Dk[i][i - 1] * x[i - 1];
      x[i] = s0 + s1;
    }
#pragma unroll
    for (int i = 0; i < kNB; i++) Li[i][lane] = x[i];
  }
  __syncthreads();
  if (slab == 0) {
    for (int e = tid; e < kNB * kNB; e += 256) {
      const int i = e >> 5, j = e & 31;
      if (i < nb && j <= i) S[(size_t)(k0 + i) * d + k0 + j] = Dk[i][j];
    }
  }
  // rows below the diagonal block: y = a L^-T, i.e. y_j = sum_{m<=j} a_m Linv[j][m] with the explicit inverse of
  // the unit lower-triangular block -> fully parallel, coalesced row access
  for (int sl = slab; sl < nslabs; sl += nslab_stride) {
    const int r0 = k0 + nb + sl * 64;
    if (sl != slab) {
      __syncthreads();
      for (int e = tid; e < 64 * kNB; e += 256) {
        const int rr = e >> 5, j = e & 31;
        As[rr][j] = (r0 + rr < d && j < nb) ? S[(size_t)(r0 + rr) * d + k0 + j] : 0.0;
      }
      __syncthreads();
    }
    for (int e = tid; e < 64 * kNB; e += 256) {
      const int rr = e >> 5, j = e & 31;
      if (r0 + rr >= d || j >= nb) continue;
      double y = 0.0;
      for (int m = 0; m <= j; m++) y += As[rr][m] * Li[j][m];
      W[(size_t)(r0 + rr) * kNB + j] = y;
      S[(size_t)(r0 + rr) * d + k0 + j] = Dk[j][j] != 0.0 ? y / Dk[j][j] : 0.0;
    }
  }
}

__global__ void __launch_bounds__(256)
k_ldlt_panel(int d, int k0, int nb, double* __restrict__ S, double* __restrict__ W, int32_t* __restrict__ flags,
             int ncheck) {
  __shared__ double sm[kLdltSmemDoubles];
  ldlt_panel_body(d, k0, nb, S, W, flags, blockIdx.x, gridDim.x, gridDim.x, sm, ncheck);
}

__device__ __forceinline__ void ldlt_update_body(int d, int k0, int nb, double* __restrict__ S,
                                                 const double* __restrict__ W, int pair_in, double* __restrict__ sm) {
  // lower-triangular tile pair (I >= J) of the trailing matrix starting at k1 = k0 + nb
  const int k1 = k0 + nb;
  int pair = pair_in, I = 0;
  while (pair > I) { pair -= I + 1; I++; }
  const int J = pair;
  const int i0 = k1 + I * 64, j0 = k1 + J * 64;
  double (*Yt)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(sm);
  double (*Lt)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(sm + 64 * (kNB + 1));
  const int tid = threadIdx.x;
  __syncthreads();
  for (int e = tid; e < 64 * kNB; e += 256) {
    const int rr = e / kNB, m = e % kNB;
    const int gi = i0 + rr, gj = j0 + rr;
    Yt[rr][m] = (gi < d && m < nb) ? W[(size_t)gi * kNB + m] : 0.0;
    Lt[rr][m] = (gj < d && m < nb) ? S[(size_t)gj * d + k0 + m] : 0.0;
  }
  __syncthreads();
  const int ti = tid >> 4, tj = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
  for (int m = 0; m < kNB; m++) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { av[a] = Yt[ti * 4 + a][m]; bv[a] = Lt[tj * 4 + a][m]; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] += av[a] * bv[b];
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int gi = i0 + ti * 4 + a, gj = j0 + tj * 4 + b;
      if (gi < d && gj <= gi) S[(size_t)gi * d + gj] -= acc[a][b];
    }
}

__global__ void __launch_bounds__(256)
k_ldlt_update(int d, int k0, int nb, double* __restrict__ S, const double* __restrict__ W) {
  __shared__ double sm[kLdltSmemDoubles];
  ldlt_update_body(d, k0, nb, S, W, blockIdx.x, sm);
}

// The whole factorisation in ONE cooperative launch (grid-wide barriers between the panel and update phases):
// for the small systems of this path (d = a few hundred) the per-panel launches are latency bound.
__global__ void __launch_bounds__(256)
k_ldlt_fused(int d, int ncols, double* __restrict__ S, double* __restrict__ W, int32_t* __restrict__ flags,
             long long* __restrict__ dbg) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[kLdltSmemDoubles];
  long long t_panel = 0, t_sync1 = 0, t_upd = 0, t_sync2 = 0;  // EMBA_DEBUG_TIMING: clocks of CTA 0 per phase
  // d = order of the bordered matrix, ncols = number of real pivots (the border row never becomes a panel of its own)
  for (int k0 = 0; k0 < ncols; k0 += kNB) {
    const int nb = min(kNB, d - k0);
    const int rows_below = d - (k0 + nb);
    const int nslabs = (rows_below + 63) / 64;
    const long long c0 = clock64();
    ldlt_panel_body(d, k0, nb, S, W, flags, blockIdx.x, gridDim.x, nslabs, sm, ncols);
    const long long c1 = clock64();
    grid.sync();
    const long long c2 = clock64();
    if (rows_below > 0) {
      const int ntile = (rows_below + 63) / 64;
      const int npair = ntile * (ntile + 1) / 2;
      for (int pr = blockIdx.x; pr < npair; pr += gridDim.x) ldlt_update_body(d, k0, nb, S, W, pr, sm);
    }
    const long long c3 = clock64();
    grid.sync();
    const long long c4 = clock64();
    t_panel += c1 - c0; t_sync1 += c2 - c1; t_upd += c3 - c2; t_sync2 += c4 - c3;
  }
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) { dbg[0] = t_panel; dbg[1] = t_sync1; dbg[2] = t_upd; dbg[3] = t_sync2; }
}

// x <- S^-1 rhs given the bordered factorisation: row d of L already holds z = D^-1 L^-1 rhs, so only the backward
// substitution L^T x = z is left (unit lower L below the diagonal, leading dimension ld = d + 1)
__global__ void __launch_bounds__(1024) k_ldlt_subst(int d, int ld, const double* __restrict__ Sm, double* __restrict__ x) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double xb[kNB];
  for (int i = tid; i < d; i += blockDim.x) x[i] = Sm[(size_t)d * ld + i];
  __syncthreads();
  // backward: L^T z = y, blocks from the end
  for (int k0 = ((d - 1) / kNB) * kNB; k0 >= 0; k0 -= kNB) {
    const int nb = min(kNB, d - k0);
    if (warp == 0) {
      // column `lane` of the block's L rows, all loads in flight at once (the serial loop below then runs from
      // registers instead of paying one L2 round trip per step)
      double lcol[kNB];
#pragma unroll
      for (int j = 0; j < kNB; j++) lcol[j] = (j < nb && lane < j) ? Sm[(size_t)(k0 + j) * ld + k0 + lane] : 0.0;
      double v = lane < nb ? x[k0 + lane] : 0.0;
#pragma unroll
      for (int j = kNB - 1; j >= 0; j--) {
        const double vj = __shfl_sync(0xffffffffu, v, j);
        v -= lcol[j] * vj;
      }
      if (lane < nb) { x[k0 + lane] = v; xb[lane] = v; }
    }
    __syncthreads();
    // x_i -= sum_{k in block} L_ki x_k for i < k0 (row k of L is contiguous in i)
    for (int i = tid; i < k0; i += blockDim.x) {
      double s = 0.0;
      for (int kk = 0; kk < nb; kk++) s += Sm[(size_t)(k0 + kk) * ld + i] * xb[kk];
      x[i] -= s;
    }
    __syncthreads();
  }
}

// x2_a = C_a (b2_a - U_a^T x1) (model.cpp:791); x1full has 3n entries (zeros for a fixed first pose)
__global__ void k_solve_x2(int64_t Np, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
                           const int64_t* __restrict__ stripoff, const double* __restrict__ strip,
                           const double* __restrict__ C, const double* __restrict__ b2,
                           const double* __restrict__ x1full, double* __restrict__ x2, int64_t own0, int64_t own1) {
  // 8 lanes per pixel walk its strip as (column 0, column 1) pairs: pair e multiplies x1[3*lo + e]; coalesced
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t a = gt >> 3;
  const int sub = (int)(gt & 7);
  double t0 = 0.0, t1 = 0.0;
  int lo = 0, hi = -1;
  if (a < Np) {
    lo = winlo[a]; hi = winhi[a];
    if (hi >= lo) {
      const double2* sp = reinterpret_cast<const double2*>(strip) + stripoff[a] * 3;
      const double* xv = x1full + 3 * lo;
      const int ne = 3 * (hi - lo + 1);
      for (int e = sub; e < ne; e += 8) {
        const double2 u = sp[e];
        const double xe = xv[e];
        t0 += u.x * xe;
        t1 += u.y * xe;
      }
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
  }
  if (a >= Np || sub != 0) return;
  if (a < own0 || a >= own1) {  // pixel owned by another rank (multi-GPU): its owner computes x2, combined by all-reduce
    x2[2 * a] = 0.0;
    x2[2 * a + 1] = 0.0;
    return;
  }
  // (an owned pixel with an empty pose window -- thres_valid_pixel <= 0 -- still gets x2 = C b2, model.cpp:791)
  const double r0 = b2[2 * a] - t0, r1 = b2[2 * a + 1] - t1;
  const double c00 = C[3 * a], c01 = C[3 * a + 1], c11 = C[3 * a + 2];
  x2[2 * a] = c00 * r0 + c01 * r1;
  x2[2 * a + 1] = c01 * r0 + c11 * r1;
}

__global__ void k_expand_x1(int n, int fix, const double* __restrict__ rhs_x, double* __restrict__ x1full) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  x1full[i] = i < 3 * fix ? 0.0 : rhs_x[i - 3 * fix];
}

// ---------------------------------------------------------------------------------------------------
// Jacobi-PCG on [[A11m, A12], [A12^T, A22m]] (model.cpp:794-840). Vectors are laid out [x1 (d) | x2 (2Np)].
// ---------------------------------------------------------------------------------------------------
constexpr int kCgChunks = 592;  // 4 CTAs per SM
// SpMV with the strips. One warp per pixel (pixels dealt round-robin to the warps of a CTA inside a contiguous
// chunk): y2_a = U_a^T p1 + A22m_a p2_a by a warp reduction, and the pixel's contribution U_a p2_a to y1 is added
// into a per-WARP accumulator in shared memory (lanes own distinct rows, pixels come in a fixed order), so the
// summation order is fixed: warp accumulators -> CTA partial (fixed order) -> k_cg_y1 (fixed order over chunks).
__device__ __forceinline__ void
cg_pix_body(double* y1s, int chunk, int nchunks, int64_t Np, int d, int fix, int n, int nwarps, const int32_t* __restrict__ winlo,
            const int32_t* __restrict__ winhi,
            const int64_t* __restrict__ stripoff, const double* __restrict__ strip, const double* __restrict__ A22,
            const double* __restrict__ A11, double lambda, const double* __restrict__ p, double* __restrict__ y,
            double* __restrict__ part, int64_t own0, int64_t own1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // y1 = A11m p1: one warp per row, rows dealt round-robin over the grid (A11 == nullptr on the ranks that do not
  // hold the pose block: their rows start from zero)
  for (int i = chunk * 8 + warp; i < d; i += nchunks * 8) {
    double s = 0.0;
    if (A11) {
      const double* row = A11 + (size_t)(3 * fix + i) * (3 * n) + 3 * fix;
      for (int j = lane; j < d; j += 32) {
        double a = row[j];
        if (j == i) a += lambda * a;
        s += a * p[j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    }
    if (lane == 0) y[i] = s;
  }
  // chunks of the pixel range this rank owns (the map rows of the other pixels were zeroed by the caller)
  const int64_t a0 = own0 + (own1 - own0) * chunk / nchunks, a1 = own0 + (own1 - own0) * (chunk + 1) / nchunks;
  for (int i = threadIdx.x; i < nwarps * d; i += blockDim.x) y1s[i] = 0.0;
  __syncthreads();
  const double* p1 = p;
  const double* p2 = p + d;
  if (warp < nwarps) {
    double* acc = y1s + (size_t)warp * d;
    // the per-pixel header (window, strip offset, p2, A22) of the NEXT pixel is requested before the current pixel's
    // rows are walked, and the rows of a pixel are requested all at once (up to 4 x 32) before any is used: a warp
    // walks ~15 pixels one after the other, and with everything loaded where it was used a C2 product took ~90 us
    // for 113 MB of strips
    int64_t a = a0 + warp;
    int lo = 0, hi = -1;
    int64_t so = 0;
    double pa = 0.0, pb = 0.0, xx = 0.0, xy = 0.0, yy = 0.0;
    auto header = [&](int64_t q) {
      lo = winlo[q]; hi = winhi[q]; so = stripoff[q];
      pa = p2[2 * q]; pb = p2[2 * q + 1];
      xx = A22[3 * q]; xy = A22[3 * q + 1]; yy = A22[3 * q + 2];
    };
    if (a < a1) header(a);
    while (a < a1) {
      const int c_lo = lo, c_hi = hi;
      const double c_pa = pa, c_pb = pb, c_xx = xx, c_xy = xy, c_yy = yy;
      const double2* sp = reinterpret_cast<const double2*>(strip + so * 6);
      const int64_t an = a + nwarps;
      if (an < a1) header(an);
      const int rows = c_hi >= c_lo ? (c_hi - c_lo + 1) * 3 : 0;
      const int g0 = 3 * (c_lo - fix);  // row of the strip's first entry in the reduced system (negative: fixed pose)
      double t0 = 0.0, t1 = 0.0;
      if (rows <= 128) {
        double2 u[4];
        double xv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int rr = lane + 32 * k;
          const bool ok = rr < rows && g0 + rr >= 0;
          u[k] = ok ? sp[rr] : make_double2(0.0, 0.0);
          xv[k] = ok ? p1[g0 + rr] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int rr = lane + 32 * k;
          if (rr < rows && g0 + rr >= 0) {
            acc[g0 + rr] += u[k].x * c_pa + u[k].y * c_pb;
            t0 += u[k].x * xv[k];
            t1 += u[k].y * xv[k];
          }
        }
      } else {
        for (int rr = lane; rr < rows; rr += 32) {
          const int grow = g0 + rr;
          if (grow < 0) continue;
          const double2 uu = sp[rr];
          acc[grow] += uu.x * c_pa + uu.y * c_pb;
          const double x = p1[grow];
          t0 += uu.x * x;
          t1 += uu.y * x;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_down_sync(0xffffffffu, t0, o);
        t1 += __shfl_down_sync(0xffffffffu, t1, o);
      }
      if (lane == 0) {
        // several GPUs: A22 is replicated, so only the pixel's owner adds the A22m term (the strips of the other
        // pixels are empty here and t0 = t1 = 0); the partial vectors are summed by one all-reduce
        y[d + 2 * a] = t0 + (c_xx + lambda * c_xx) * c_pa + c_xy * c_pb;
        y[d + 2 * a + 1] = t1 + c_xy * c_pa + (c_yy + lambda * c_yy) * c_pb;
      }
      __syncwarp();
      a = an;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nwarps; w++) s += y1s[(size_t)w * d + i];
    part[(size_t)chunk * d + i] = s;
  }
}

__global__ void __launch_bounds__(256)
k_cg_pix(int64_t Np, int d, int fix, int n, int nwarps, const int32_t* __restrict__ winlo, const int32_t* __restrict__ winhi,
         const int64_t* __restrict__ stripoff, const double* __restrict__ strip, const double* __restrict__ A22,
         const double* __restrict__ A11, double lambda, const double* __restrict__ p, double* __restrict__ y,
         double* __restrict__ part, int64_t own0, int64_t own1, const int* __restrict__ done) {
  extern __shared__ double y1s[];  // [nwarps][d]
  if (*done) return;
  cg_pix_body(y1s, blockIdx.x, gridDim.x, Np, d, fix, n, nwarps, winlo, winhi, stripoff, strip, A22, A11, lambda, p, y, part,
              own0, own1);
}

// y1 += sum over the chunks' partial rows: one WARP per row (lane l takes chunks l, l + 32, ... then a fixed shuffle
// tree), so the ~600 loads of a row are independent and in flight together; a thread walking them one after the other
// was 2/3 of a PCG iteration on C2
__global__ void __launch_bounds__(256) k_cg_y1(int d, int chunks, const double* __restrict__ part, double* __restrict__ y,
                                               const int* __restrict__ done) {
  if (*done) return;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= d) return;
  double s = 0.0;
  for (int c = lane; c < chunks; c += 32) s += part[(size_t)c * d + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[i] += s;
}

__global__ void k_mul(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] * b[i];
}
// rhs vector b = [b1(fixed) | b2] and the inverse Jacobi diagonal (Eigen DiagonalPreconditioner: 1/diag, 1 if 0)
__global__ void k_cg_setup(int d, int fix, int n, int64_t Np, const double* __restrict__ A11,
                           const double* __restrict__ b1, const double* __restrict__ A22,
                           const double* __restrict__ b2, double lambda, double* __restrict__ b,
                           double* __restrict__ invd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = d + 2 * Np;
  if (i >= tot) return;
  double dg, bv;
  if (i < d) {
    const double a = A11[(size_t)(3 * fix + i) * (3 * n) + 3 * fix + i];
    dg = a + lambda * a;
    bv = b1[3 * fix + i];
  } else {
    const int64_t k = i - d, a = k >> 1;
    const double v = (k & 1) ? A22[3 * a + 2] : A22[3 * a];
    dg = v + lambda * v;
    bv = b2[k];
  }
  b[i] = bv;
  invd[i] = dg != 0.0 ? 1.0 / dg : 1.0;
}

// ---------------------------------------------------------------------------------------------------
// state update
__global__ void k_update_quat(int n, int fix, const double* __restrict__ qin, const double* __restrict__ x1full,
                              double* __restrict__ qout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 q = reinterpret_cast<const double4*>(qin)[i];
  double4 r = q;
  if (i >= fix) {
    // R_i <- Exp(dphi_i) * R_i, left perturbation (src/utils/trajectory.cpp:296-304); Sophus renormalises
    const Vec3 w = {x1full[3 * i], x1full[3 * i + 1], x1full[3 * i + 2]};
    r = quat_mul(so3_exp(w), q);
    const double nrm = sqrt(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    r.x /= nrm; r.y /= nrm; r.z /= nrm; r.w /= nrm;
  }
  reinterpret_cast<double4*>(qout)[i] = r;
}

// active pixels += damping * x2, inactive pixels <- 0 (model.cpp:863-903)
__global__ void k_update_map(int64_t P, const int32_t* __restrict__ amap, const double* __restrict__ x2,
                             double damping, const double* __restrict__ Gx, const double* __restrict__ Gy,
                             double* __restrict__ Gxo, double* __restrict__ Gyo) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int32_t a = amap[p];
  if (a >= 0) {
    Gxo[p] = Gx[p] + damping * x2[2 * (int64_t)a];
    Gyo[p] = Gy[p] + damping * x2[2 * (int64_t)a + 1];
  } else {
    Gxo[p] = 0.0;
    Gyo[p] = 0.0;
  }
}

int solve_schur(Handle* h, double lambda, int fix) {
  static const bool dbg = getenv("EMBA_DEBUG_TIMING") != nullptr;
  cudaEvent_t de[6];
  if (dbg) { for (auto& e : de) cudaEventCreate(&e); cudaEventRecord(de[0], h->stream); }
  const int n = h->n;
  const int d = 3 * (n - fix);
  const int64_t Np = h->Np;
  const int T = 256;
  if (Np > 0) { k_a22_inv<<<ceil_div64(Np, T), T, 0, h->stream>>>(Np, h->d_A22, lambda, h->d_C); EMBA_LAUNCH_CHECK(); }
  const int nt = (d + 1 + kST - 1) / kST;
  const int npairs = nt * (nt + 1) / 2;
  static const int zmult = getenv("EMBA_SCHUR_ZMULT") ? atoi(getenv("EMBA_SCHUR_ZMULT")) : 16;
  const int64_t own0 = h->world > 1 ? Np * h->rank / h->world : 0, own1 = h->world > 1 ? Np * (h->rank + 1) / h->world : Np;
  int Z = std::max(1, (h->sm_count * zmult) / npairs);
  if (!getenv("EMBA_SCHUR_ZMULT") && 48.0 * (double)h->sv_strip_total / 1e6 > 96.0)
    Z = std::max(Z, std::min(160, (int)(48.0 * (double)h->sv_strip_total / 16e6) + 1));
  Z = (int)std::min<int64_t>(Z, std::max<int64_t>(1, (own1 - own0 + kSchurThreads - 1) / kSchurThreads));
  EMBA_TRY(dev_reserve(h, &h->d_Spart, &h->Spart_cap, (int64_t)Z * npairs * kST * kST));
  EMBA_TRY(fill_strip_masks(h));  // occupancy masks of the strips k_pix left in global memory (once per assembly)
  // strips beyond the L2's reach: tile-pair-fastest CTA order with pixel chunks of ~16 MB of strips, so the chunk a
  // wave of CTAs shares is read from HBM once instead of once per tile pair (C4: 4.67 -> 4.0 ms, C3: 5.5 -> 5.0 ms;
  // small problems keep the heaviest-pair-first order, which is 6 % faster on C2)
  static const int order_env = getenv("EMBA_SCHUR_ORDER") ? atoi(getenv("EMBA_SCHUR_ORDER")) : -1;
  const double strip_mb = 48.0 * (double)h->sv_strip_total / 1e6;
  const int order = order_env >= 0 ? order_env : (strip_mb > 96.0 ? 1 : 0);
  dim3 grid(Z, npairs);
  if (order) grid = dim3((unsigned)((int64_t)Z * npairs), 1);
  // resident CTAs per SM the tile kernel is compiled for (registers 108 / 92 / 78 / 72, no spills). The kernel waits on
  // fixed-latency dependencies (smem load -> shuffle -> scale -> DMMA), so more warps pay: whole solve on C4 / C3 / C2
  // (ms) 4: 5.28 / 5.04 / 0.826, 5: 4.85 / 4.63 / 0.797, 6: 4.65 / 4.45 / 0.777, 7: 4.77 / 4.57 / 0.82. Read per call.
  const int schur_occ = getenv("EMBA_SCHUR_OCC") ? atoi(getenv("EMBA_SCHUR_OCC")) : 6;
  if (schur_occ >= 7) k_schur_tiles<7><<<grid, kSchurThreads, 0, h->stream>>>(own0, own1, d, fix, nt, Z, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip,
                                             h->d_C, h->d_b2, h->sv_gmask, h->pose_group, h->d_Spart, order);
  else if (schur_occ == 6) k_schur_tiles<6><<<grid, kSchurThreads, 0, h->stream>>>(own0, own1, d, fix, nt, Z, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip,
                                             h->d_C, h->d_b2, h->sv_gmask, h->pose_group, h->d_Spart, order);
  else if (schur_occ == 5) k_schur_tiles<5><<<grid, kSchurThreads, 0, h->stream>>>(own0, own1, d, fix, nt, Z, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip,
                                             h->d_C, h->d_b2, h->sv_gmask, h->pose_group, h->d_Spart, order);
  else k_schur_tiles<4><<<grid, kSchurThreads, 0, h->stream>>>(own0, own1, d, fix, nt, Z, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip,
                                             h->d_C, h->d_b2, h->sv_gmask, h->pose_group, h->d_Spart, order);
  EMBA_LAUNCH_CHECK();
  const int64_t tot = (int64_t)d * (d + 1);
  if (dbg) cudaEventRecord(de[1], h->stream);
  if (h->world == 1) {
    k_schur_finish<<<ceil_div64(tot, T), T, 0, h->stream>>>(d, fix, n, nt, npairs, Z, h->d_Spart, h->d_A11, h->d_b1,
                                                           lambda, h->d_S, h->d_rhs, nullptr, 0);
    EMBA_LAUNCH_CHECK();
  } else {
    // every rank holds the Schur contributions of the pixels it owns (and, right after an assembly, only its own
    // partial A11 / b1): form the local S and rhs, combine them over NVLink with ONE all-reduce.
    // A11m = A11 + lambda*diag(A11) is linear in A11, so summing the per-rank damped partials is exact.
    if (h->a11_partial) {
      k_schur_finish<<<ceil_div64(tot, T), T, 0, h->stream>>>(d, fix, n, nt, npairs, Z, h->d_Spart, h->d_A11, h->d_b1,
                                                             lambda, h->d_S, h->d_rhs, nullptr, 0);
      EMBA_LAUNCH_CHECK();
      EMBA_TRY(comm_allreduce(h, h->d_S, (int64_t)(d + 1) * (d + 1), 1));  // bordered: the rhs row rides along
    } else {
      // A11 / b1 already combined (after emba_get_normal_eq): all-reduce only the Schur sums
      EMBA_TRY(dev_reserve(h, &h->d_cg, &h->cg_cap, tot));
      k_schur_finish<<<ceil_div64(tot, T), T, 0, h->stream>>>(d, fix, n, nt, npairs, Z, h->d_Spart, h->d_A11, h->d_b1,
                                                             lambda, h->d_S, h->d_rhs, h->d_cg, 1);
      EMBA_LAUNCH_CHECK();
      EMBA_TRY(comm_allreduce(h, h->d_cg, tot, 1));
      k_schur_finish<<<ceil_div64(tot, T), T, 0, h->stream>>>(d, fix, n, nt, npairs, Z, h->d_Spart, h->d_A11, h->d_b1,
                                                             lambda, h->d_S, h->d_rhs, h->d_cg, 2);
      EMBA_LAUNCH_CHECK();
    }
  }
  if (dbg) cudaEventRecord(de[2], h->stream);
  EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
  const int db = d + 1;  // order of the bordered matrix
  EMBA_TRY(dev_reserve(h, &h->d_ldlt_w, &h->ldlt_w_cap, (int64_t)db * kNB));
  if (db <= 1536) {
    // small system: one cooperative launch, grid barriers instead of 2 launches per panel
    int per_sm = 0;
    EMBA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ldlt_fused, 256, 0));
    const int ntile0 = (db + 63) / 64;
    int grid_f = std::min(h->sm_count * std::max(1, per_sm), std::max(1, ntile0 * (ntile0 + 1) / 2));
    int dd = db, ncols = d;
    double* Sp = h->d_S;
    double* Wp = h->d_ldlt_w;
    int32_t* fp = h->d_flags;
    static long long* d_dbg = nullptr;
    if (dbg && !d_dbg) cudaMalloc((void**)&d_dbg, 8 * sizeof(long long));
    long long* dp = dbg ? d_dbg : nullptr;
    void* args[] = {&dd, &ncols, &Sp, &Wp, &fp, &dp};
    EMBA_CUDA(cudaLaunchCooperativeKernel((void*)k_ldlt_fused, dim3(grid_f), dim3(256), args, 0, h->stream));
    h->launches++;
    if (dbg) {
      long long hc[4];
      cudaStreamSynchronize(h->stream);
      cudaMemcpy(hc, d_dbg, sizeof(hc), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[emba ldlt] grid %d clocks(CTA 0): panel %lld sync %lld update %lld sync %lld\n", grid_f, hc[0], hc[1], hc[2], hc[3]);
    }
  } else {
    for (int k0 = 0; k0 < d; k0 += kNB) {
      const int nb = std::min(kNB, db - k0);
      const int rows_below = db - (k0 + nb);
      const int slabs = std::max(1, (rows_below + 63) / 64);
      k_ldlt_panel<<<slabs, 256, 0, h->stream>>>(db, k0, nb, h->d_S, h->d_ldlt_w, h->d_flags, d);
      h->launches++;
      if (rows_below > 0) {
        const int nt = (rows_below + 63) / 64;
        k_ldlt_update<<<nt * (nt + 1) / 2, 256, 0, h->stream>>>(db, k0, nb, h->d_S, h->d_ldlt_w);
        h->launches++;
      }
    }
  }
  if (dbg) cudaEventRecord(de[3], h->stream);
  k_ldlt_subst<<<1, 1024, 0, h->stream>>>(d, db, h->d_S, h->d_rhs);
  EMBA_LAUNCH_CHECK();
  if (dbg) cudaEventRecord(de[4], h->stream);
  k_expand_x1<<<ceil_div64(3 * n, T), T, 0, h->stream>>>(n, fix, h->d_rhs, h->d_x1);
  EMBA_LAUNCH_CHECK();
  if (Np > 0) {
    k_solve_x2<<<ceil_div64(Np * 8, 256), 256, 0, h->stream>>>(Np, h->sv_winlo, h->sv_winhi, h->sv_stripoff, h->sv_strip,
                                                           h->d_C, h->d_b2, h->d_x1, h->d_x2, own0, own1);
    EMBA_LAUNCH_CHECK();
    if (h->world > 1) EMBA_TRY(comm_allreduce(h, h->d_x2, 2 * Np, 1));  // owners contribute, the others hold zeros
  }
  int32_t fl = 0;
  if (dbg) cudaEventRecord(de[5], h->stream);
  EMBA_CUDA(cudaMemcpyAsync(&fl, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  if (dbg) {
    float a, b, c, e2, f;
    cudaEventElapsedTime(&a, de[0], de[1]); cudaEventElapsedTime(&b, de[1], de[2]); cudaEventElapsedTime(&c, de[2], de[3]);
    cudaEventElapsedTime(&e2, de[3], de[4]); cudaEventElapsedTime(&f, de[4], de[5]);
    fprintf(stderr, "[emba solve] d=%d a22inv+tiles %.3f finish %.3f ldlt %.3f subst %.3f x2 %.3f ms\n", d, a, b, c, e2, f);
    for (auto& e : de) cudaEventDestroy(e);
  }
  if (fl & 8) { h->err = "non-finite pivot in the LDL^T factorisation of the Schur complement"; return EMBA_E_NUMERIC; }
  return EMBA_OK;
}

// ---- device-resident CG control: the scalars of Eigen's loop (ConjugateGradient.h:26-96) live in one small
// struct on the device. An iteration is FIVE launches (product, dot, and the three vector updates); every kernel sums
// the previous kernel's per-block partials itself (all blocks in the same fixed order, so they agree bit for bit),
// reads the stopping flag a PREVIOUS kernel wrote and returns at once when it has fired. The host only enqueues
// iterations (in chunks) and looks at the flag once per chunk.
struct CgScal {
  double rhs2, thr, absNew[2], rn2;
  int done, iters;
};
constexpr int kCgDotGrid = 296;

__device__ __forceinline__ double cg_sum_partials(const double* __restrict__ part, double* sh) {
  // fixed-order sum of the kCgDotGrid partials, by every block alike
  double s = 0;
  for (int b = threadIdx.x; b < kCgDotGrid; b += 256) s += part[b];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  const double v = sh[0];
  __syncthreads();
  return v;
}
__device__ __forceinline__ void cg_block_partial(double s, double* sh, double* __restrict__ part) {
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__device__ __forceinline__ void cg_block_partial_at(double s, double* sh, double* __restrict__ part, int vb) {
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[vb] = sh[0];
  __syncthreads();
}

// rhsNorm2 == 0 -> x = 0, 0 iterations; threshold = max(tol^2 rhsNorm2, smallest normal); residual = rhs (x0 = 0);
// absNew = r . (M^-1 r)
__global__ void __launch_bounds__(256) k_cg_init(const double* __restrict__ part_bb, const double* __restrict__ part_rp,
                                                 CgScal* sc, double tol) {
  __shared__ double sh[256];
  const double rhs2 = cg_sum_partials(part_bb, sh);
  const double absNew = cg_sum_partials(part_rp, sh);
  if (threadIdx.x == 0) {
    sc->rhs2 = rhs2;
    sc->thr = fmax(tol * tol * rhs2, 2.2250738585072014e-308);
    sc->rn2 = rhs2;
    sc->absNew[0] = absNew;
    sc->iters = 0;
    sc->done = (rhs2 == 0.0 || rhs2 < sc->thr) ? 1 : 0;
  }
}

// partials of a . b (fixed grid-stride order)
__global__ void __launch_bounds__(256) k_cg_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                double* __restrict__ part, const CgScal* sc, int check) {
  __shared__ double sh[256];
  if (check && sc->done) return;
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += a[i] * b[i];
  cg_block_partial(s, sh, part);
}
// x += alpha p, r -= alpha tmp, partials of r.r   (alpha = absNew / p.tmp)
__global__ void __launch_bounds__(256) k_cg_step1(int64_t n, const double* __restrict__ p, const double* __restrict__ tmp,
                                                  double* __restrict__ x, double* __restrict__ r,
                                                  const double* __restrict__ part_in, double* __restrict__ part_out,
                                                  const CgScal* sc, int par) {
  __shared__ double sh[256];
  if (sc->done) return;
  const double alpha = sc->absNew[par] / cg_sum_partials(part_in, sh);
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    x[i] += alpha * p[i];
    const double ri = r[i] + (-alpha) * tmp[i];
    r[i] = ri;
    s += ri * ri;
  }
  cg_block_partial(s, sh, part_out);
}
// stopping test on r.r (`break` before i++, ConjugateGradient.h:78-79); z = invd .* r, partials of r.z
__global__ void __launch_bounds__(256) k_cg_step2(int64_t n, const double* __restrict__ invd, const double* __restrict__ r,
                                                  double* __restrict__ z, const double* __restrict__ part_in,
                                                  double* __restrict__ part_out, CgScal* sc) {
  __shared__ double sh[256];
  if (sc->done) return;
  const double rn2 = cg_sum_partials(part_in, sh);
  const bool stop = rn2 < sc->thr;
  if (blockIdx.x == 0 && threadIdx.x == 0) { sc->rn2 = rn2; if (stop) sc->done = 1; }  // read by LATER kernels only
  if (stop) return;
  double s = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double zi = invd[i] * r[i];
    z[i] = zi;
    s += r[i] * zi;
  }
  cg_block_partial(s, sh, part_out);
}
// beta = absNew' / absNew; p = z + beta p; i++ ; while (i < maxIters)
__global__ void __launch_bounds__(256) k_cg_step3(int64_t n, const double* __restrict__ z, double* __restrict__ p,
                                                  const double* __restrict__ part_in, CgScal* sc, int par, int max_iter) {
  __shared__ double sh[256];
  if (sc->done) return;
  const double absNew = cg_sum_partials(part_in, sh);
  const double beta = absNew / sc->absNew[par];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) p[i] = z[i] + beta * p[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->absNew[par ^ 1] = absNew;  // the other parity: nobody reads it in this launch
    sc->iters += 1;
  }
}
// the iteration bound is applied by the first kernel of the NEXT iteration's chain (so that `done` is never written and
// read inside one launch): folded into the product kernel's guard below
__global__ void k_cg_bound(CgScal* sc, int max_iter) {
  if (!sc->done && sc->iters >= max_iter) sc->done = 1;
}

// ---- the whole iteration loop as ONE cooperative kernel (single GPU): the six phases of an iteration are separated
// by grid barriers instead of kernel boundaries. Same grids in spirit -- kCgChunks CTAs walk the pixel chunks, the first
// kCgDotGrid of them do the vector phases -- and the same per-block partials summed in the same fixed order, so the
// arithmetic, the stopping decision and the iteration count are bit for bit those of the multi-launch chain above
// (which stays for several GPUs, where an all-reduce sits inside the product). Every block derives the stopping test
// from the same partials, so all blocks leave the loop together.
struct CgArgs {
  int64_t Np, tot;
  int d, fix, n, nwarps, max_iter;
  const int32_t *winlo, *winhi;
  const int64_t* stripoff;
  const double *strip, *A22, *A11, *invd;
  double lambda;
  double *x, *r, *p, *z, *tmp, *ypart, *pa, *pb, *pc;
  CgScal* sc;
};

__global__ void __launch_bounds__(256, 4) k_cg_persist(CgArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double y1s[];
  __shared__ double sh[256];
  const int bid = blockIdx.x;
  const int64_t tot = a.tot;
  CgScal* sc = a.sc;
  if (sc->done) return;  // written by k_cg_init, an earlier launch: uniform
  const double thr = sc->thr;
  const int nrowblk = (a.d + 7) / 8;
  for (int k = 0;; k++) {
    const int par = k & 1;
    // tmp = A p
    cg_pix_body(y1s, bid, gridDim.x, a.Np, a.d, a.fix, a.n, a.nwarps, a.winlo, a.winhi, a.stripoff, a.strip, a.A22, a.A11,
                a.lambda, a.p, a.tmp, a.ypart, 0, a.Np);
    grid.sync();
    for (int vb = bid; vb < nrowblk; vb += gridDim.x) {  // k_cg_y1
      const int i = vb * 8 + (threadIdx.x >> 5);
      const int lane = threadIdx.x & 31;
      if (i < a.d) {
        double s = 0.0;
        for (int c = lane; c < (int)gridDim.x; c += 32) s += a.ypart[(size_t)c * a.d + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) a.tmp[i] += s;
      }
    }
    grid.sync();
    if (bid < kCgDotGrid) {  // k_cg_dot: partials of p . tmp
      double s = 0;
      for (int64_t i = (int64_t)bid * 256 + threadIdx.x; i < tot; i += (int64_t)kCgDotGrid * 256) s += a.p[i] * a.tmp[i];
      cg_block_partial_at(s, sh, a.pa, bid);
    }
    grid.sync();
    if (bid < kCgDotGrid) {  // k_cg_step1
      const double alpha = sc->absNew[par] / cg_sum_partials(a.pa, sh);
      double s = 0;
      for (int64_t i = (int64_t)bid * 256 + threadIdx.x; i < tot; i += (int64_t)kCgDotGrid * 256) {
        a.x[i] += alpha * a.p[i];
        const double ri = a.r[i] + (-alpha) * a.tmp[i];
        a.r[i] = ri;
        s += ri * ri;
      }
      cg_block_partial_at(s, sh, a.pb, bid);
    }
    grid.sync();
    // k_cg_step2: every block takes the stopping decision from the same partials
    const double rn2 = cg_sum_partials(a.pb, sh);
    const bool stop = rn2 < thr;
    if (bid == 0 && threadIdx.x == 0) { sc->rn2 = rn2; if (stop) sc->done = 1; }
    if (stop) break;
    if (bid < kCgDotGrid) {
      double s = 0;
      for (int64_t i = (int64_t)bid * 256 + threadIdx.x; i < tot; i += (int64_t)kCgDotGrid * 256) {
        const double zi = a.invd[i] * a.r[i];
        a.z[i] = zi;
        s += a.r[i] * zi;
      }
      cg_block_partial_at(s, sh, a.pc, bid);
    }
    grid.sync();
    if (bid < kCgDotGrid) {  // k_cg_step3
      const double absNew = cg_sum_partials(a.pc, sh);
      const double beta = absNew / sc->absNew[par];
      for (int64_t i = (int64_t)bid * 256 + threadIdx.x; i < tot; i += (int64_t)kCgDotGrid * 256) a.p[i] = a.z[i] + beta * a.p[i];
      if (bid == 0 && threadIdx.x == 0) {
        sc->absNew[par ^ 1] = absNew;
        sc->iters += 1;
      }
    }
    if (k + 1 >= a.max_iter) {  // k_cg_bound
      if (bid == 0 && threadIdx.x == 0) sc->done = 1;
      break;
    }
    grid.sync();
  }
}

int solve_pcg(Handle* h, double lambda, int fix, int* iters_out, double* err_out) {
  const int n = h->n;
  const int d = 3 * (n - fix);
  const int64_t Np = h->Np;
  // Several GPUs: every vector is replicated; the matrix is not. A12 lives with the pixel owners (solve view), A11 /
  // b1 are combined here once if they are still per-rank partials, A22 / b2 are already global. A product y = A v
  // is then a sum of per-rank partial vectors -- rank 0 adds the A11m block, owners add their A22m blocks and their
  // strips' contributions -- combined by ONE all-reduce of (d + 2 Np) doubles per iteration, issued from the device
  // timeline like every other step. All ranks see the same y, so the scalars need no communication and every rank
  // takes the same decisions.
  const int W = h->world;
  if (W > 1 && h->a11_partial) {
    EMBA_TRY(comm_allreduce(h, h->d_A11, (int64_t)9 * n * n, 1));
    EMBA_TRY(comm_allreduce(h, h->d_b1, (int64_t)3 * n, 1));
    h->a11_partial = false;
  }
  const int64_t own0 = W > 1 ? Np * h->rank / W : 0, own1 = W > 1 ? Np * (h->rank + 1) / W : Np;
  const int64_t tot = d + 2 * Np;
  const int T = 256;
  const int G = ceil_div64(tot, T);
  const int max_iter = 100;
  const double tol = 1e-6;  // model.cpp:828-831
  // vectors: b, invd, x, r, p, z, tmp + per-chunk y1 partials
  EMBA_TRY(dev_reserve(h, &h->d_cg, &h->cg_cap, 7 * tot + (int64_t)kCgChunks * d + 64));
  EMBA_TRY(dev_reserve(h, &h->d_part, &h->part_cap, 4096 * 4));
  double* b = h->d_cg;
  double* invd = b + tot;
  double* x = invd + tot;
  double* r = x + tot;
  double* p = r + tot;
  double* z = p + tot;
  double* tmp = z + tot;
  double* ypart = tmp + tot;
  CgScal* sc = reinterpret_cast<CgScal*>(h->d_scal + 8);
  double* pa = h->d_part;         // three partial buffers, used round robin by the chain
  double* pb = h->d_part + 1024;
  double* pc = h->d_part + 2048;
  k_cg_setup<<<G, T, 0, h->stream>>>(d, fix, n, Np, h->d_A11, h->d_b1, h->d_A22, h->d_b2, lambda, b, invd);
  EMBA_LAUNCH_CHECK();
  EMBA_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * tot, h->stream));
  EMBA_CUDA(cudaMemcpyAsync(r, b, sizeof(double) * tot, cudaMemcpyDeviceToDevice, h->stream));  // x0 = 0
  // y = A v
  auto matvec = [&](const double* v, double* y) -> int {
    const int nwarps = std::max(1, std::min(8, (int)((48 * 1024) / (sizeof(double) * (size_t)d))));
    const size_t shm = sizeof(double) * (size_t)d * nwarps;
    if (shm > 48 * 1024) EMBA_CUDA(cudaFuncSetAttribute(k_cg_pix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    if (W > 1) EMBA_CUDA(cudaMemsetAsync(y + d, 0, sizeof(double) * 2 * Np, h->stream));
    k_cg_pix<<<kCgChunks, 256, shm, h->stream>>>(Np, d, fix, n, nwarps, h->sv_winlo, h->sv_winhi, h->sv_stripoff,
                                                 h->sv_strip, h->d_A22, h->rank == 0 ? h->d_A11 : nullptr, lambda, v, y,
                                                 ypart, own0, own1, &sc->done);
    h->launches++;
    EMBA_CUDA(cudaGetLastError());
    k_cg_y1<<<ceil_div64(d, 8), 256, 0, h->stream>>>(d, kCgChunks, ypart, y, &sc->done);
    h->launches++;
    if (W > 1) EMBA_TRY(comm_allreduce(h, y, tot, 1));
    return EMBA_OK;
  };
  // rhsNorm2, stopping threshold, first search direction
  k_cg_dot<<<kCgDotGrid, 256, 0, h->stream>>>(tot, b, b, pa, sc, 0);
  k_mul<<<G, T, 0, h->stream>>>(tot, invd, r, p);
  k_cg_dot<<<kCgDotGrid, 256, 0, h->stream>>>(tot, r, p, pb, sc, 0);
  k_cg_init<<<1, 256, 0, h->stream>>>(pa, pb, sc, tol);
  h->launches += 4;
  EMBA_CUDA(cudaGetLastError());
  // one GPU, EMBA_CG_PERSIST=1: the whole loop as one cooperative launch (k_cg_persist) when kCgChunks CTAs are
  // co-resident. Measured against the chain of launches below (same arithmetic, bit-identical results, same iteration
  // counts): C2 106 vs 108 us per iteration, C4 850 vs 711 us -- six grid barriers cost what six launches cost, so the
  // launches are not what an iteration spends its time on and the chain stays the default.
  bool persisted = false;
  const bool persist_env = getenv("EMBA_CG_PERSIST") && atoi(getenv("EMBA_CG_PERSIST")) == 1;  // read per call
  if (W == 1 && persist_env && Np > 0) {
    const int nwarps = std::max(1, std::min(8, (int)((48 * 1024) / (sizeof(double) * (size_t)d))));
    const size_t shm = sizeof(double) * (size_t)d * nwarps;
    int per_sm = 0;
    if (shm <= 48 * 1024 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persist, 256, shm) == cudaSuccess &&
        per_sm * h->sm_count >= kCgChunks) {
      CgArgs a;
      a.Np = Np; a.tot = tot; a.d = d; a.fix = fix; a.n = n; a.nwarps = nwarps; a.max_iter = max_iter;
      a.winlo = h->sv_winlo; a.winhi = h->sv_winhi; a.stripoff = h->sv_stripoff; a.strip = h->sv_strip;
      a.A22 = h->d_A22; a.A11 = h->d_A11; a.invd = invd; a.lambda = lambda;
      a.x = x; a.r = r; a.p = p; a.z = z; a.tmp = tmp; a.ypart = ypart; a.pa = pa; a.pb = pb; a.pc = pc; a.sc = sc;
      void* args[] = {&a};
      EMBA_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_persist, dim3(kCgChunks), dim3(256), args, shm, h->stream));
      h->launches++;
      persisted = true;
    }
    cudaGetLastError();
  }
  // iterations: enqueued in chunks, the flag is looked at once per chunk (kernels after the stop are no-ops)
  int64_t* hflag = h->h_pin + 20;
  const int chunk = 10;
  for (int it0 = 0; it0 < max_iter && !persisted; it0 += chunk) {
    for (int k = it0; k < it0 + chunk && k < max_iter; k++) {
      EMBA_TRY(matvec(p, tmp));
      k_cg_dot<<<kCgDotGrid, 256, 0, h->stream>>>(tot, p, tmp, pa, sc, 1);
      k_cg_step1<<<kCgDotGrid, 256, 0, h->stream>>>(tot, p, tmp, x, r, pa, pb, sc, k & 1);
      k_cg_step2<<<kCgDotGrid, 256, 0, h->stream>>>(tot, invd, r, z, pb, pc, sc);
      k_cg_step3<<<kCgDotGrid, 256, 0, h->stream>>>(tot, z, p, pc, sc, k & 1, max_iter);
      h->launches += 4;
    }
    k_cg_bound<<<1, 1, 0, h->stream>>>(sc, max_iter);
    EMBA_CUDA(cudaGetLastError());
    EMBA_CUDA(cudaMemcpyAsync(hflag, &sc->done, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
    if (reinterpret_cast<int*>(hflag)[0]) break;
  }
  CgScal hs;
  EMBA_CUDA(cudaMemcpyAsync(&hs, sc, sizeof(CgScal), cudaMemcpyDeviceToHost, h->stream));
  k_expand_x1<<<ceil_div64(3 * n, T), T, 0, h->stream>>>(n, fix, x, h->d_x1);
  EMBA_LAUNCH_CHECK();
  if (Np > 0) EMBA_CUDA(cudaMemcpyAsync(h->d_x2, x + d, sizeof(double) * 2 * Np, cudaMemcpyDeviceToDevice, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  if (iters_out) *iters_out = hs.iters;
  if (err_out) *err_out = hs.rhs2 == 0.0 ? 0.0 : sqrt(hs.rn2 / hs.rhs2);
  return EMBA_OK;
}

int make_candidate(Handle* h, double damping, int fix) {
  StateSlot& c = h->st[h->cur];
  StateSlot& k = h->st[1 - h->cur];
  k_update_quat<<<ceil_div64(h->n, 128), 128, 0, h->stream>>>(h->n, fix, c.quat, h->d_x1, k.quat);
  EMBA_LAUNCH_CHECK();
  k_update_map<<<ceil_div64(h->P, 256), 256, 0, h->stream>>>(h->P, h->d_amap, h->d_x2, damping, c.Gx, c.Gy, k.Gx, k.Gy);
  EMBA_LAUNCH_CHECK();
  k.evaluated = false;
  return EMBA_OK;
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_solve(emba_handle_t hh, double lambda, int32_t use_cg, int32_t fix, double* x1_out, double* x2_out,
               int32_t* cg_iters, double* cg_error) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->formed) { h->err = "emba_solve: form the normal equations first"; return EMBA_E_ARG; }
  fix = fix ? 1 : 0;
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  if (use_cg) {
    int it = 0; double er = 0;
    EMBA_TRY(solve_pcg(h, lambda, fix, &it, &er));
    if (cg_iters) *cg_iters = it;
    if (cg_error) *cg_error = er;
  } else {
    EMBA_TRY(solve_schur(h, lambda, fix));
  }
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  h->solved = true;
  h->solved_fix = fix;
  const int d = 3 * (h->n - fix);
  if (x1_out) EMBA_CUDA(cudaMemcpyAsync(x1_out, h->d_x1 + 3 * fix, sizeof(double) * d, cudaMemcpyDeviceToHost, h->stream));
  if (x2_out && h->Np) EMBA_CUDA(cudaMemcpyAsync(x2_out, h->d_x2, sizeof(double) * 2 * h->Np, cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->t_ms[5] = ms;
  return EMBA_OK;
}

int emba_make_candidate(emba_handle_t hh, double damping, int32_t fix) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->solved) { h->err = "emba_make_candidate: solve first"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_TRY(make_candidate(h, damping, fix ? 1 : 0));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  return EMBA_OK;
}

int emba_accept_candidate(emba_handle_t hh) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (!h->st[1 - h->cur].evaluated) { h->err = "emba_accept_candidate: candidate not evaluated"; return EMBA_E_ARG; }
  h->cur = 1 - h->cur;  // solver.cpp:299-317: the candidate (state, residuals, num_ev_map) becomes current
  h->formed = false;
  h->solved = false;
  return EMBA_OK;
}

}  // extern "C"
