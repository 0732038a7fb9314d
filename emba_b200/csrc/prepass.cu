// Static structure of a time window: handle lifetime, event ingestion, batch mid-times, the per-sensor-pixel
// pair links that replace the reference's EventMap (include/emba/event_map.h:22-113), and -- once the spline
// time base is known -- the canonical measurement order (sorted by the control-pose pair a measurement
// touches) with its work items. Everything here runs once per window / per spline base, never per LM iteration.
//
// SURVEY section 8(f) N1. The per-window structures live in two arenas (event level, measurement level) plus a
// scratch arena, all grow-only: a window no larger than the previous one allocates nothing. Sorts and scans are the
// library's own (prims.cu). Events come either from host arrays (emba_set_events: pinned sources are copied
// directly, pageable ones are staged through two pinned buffers) or from a device-resident event sequence
// (emba_events_*, emba_set_events_dev: no host copy at all).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "emba_internal.cuh"

namespace emba {

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
__global__ void k_spix(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, int Ws, int Hs, int64_t N,
                       uint32_t* __restrict__ spix, uint32_t* __restrict__ key, int32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t xi = x[i], yi = y[i];
  uint32_t s = 0;
  if (xi >= (uint32_t)Ws || yi >= (uint32_t)Hs) atomicOr(flags, 1);
  else s = yi * (uint32_t)Ws + xi;
  spix[i] = s;
  key[i] = s;
}

// the two timestamps per batch the mid-time needs (device-resident sequences)
__global__ void k_batch_tpair(const int64_t* __restrict__ t, int64_t B, int64_t* __restrict__ tpair) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  tpair[2 * b] = t[b * kBatch];
  tpair[2 * b + 1] = t[b * kBatch + kBatch - 1];
}

// ros::Duration(double) rounding of the half span, as the reference evaluates `t_batch_bgn + timespan * 0.5`
// (src/emba/model.cpp:116-119): Duration::operator*(double) is Duration(toSec()*scale); Duration(double d) is
// sec = floor(d), nsec = round((d - sec)*1e9). Round-to-nearest intrinsics keep the compiler from contracting the
// products into FMAs, so odd spans round exactly like the CPU reference.
__global__ void k_batch_mid(const int64_t* __restrict__ tpair, int64_t B, int64_t* __restrict__ tmid) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t t_bgn = tpair[2 * b], span = tpair[2 * b + 1] - t_bgn;
  int64_t sec = span / 1000000000LL, nsec = span % 1000000000LL;
  if (nsec < 0) { nsec += 1000000000LL; sec -= 1; }
  const double tosec = __dadd_rn((double)sec, __dmul_rn(1e-9, (double)nsec));
  const double d = __dmul_rn(tosec, 0.5);
  const double s = floor(d);
  const double frac = __dmul_rn(__dsub_rn(d, s), 1e9);
  tmid[b] = t_bgn + (int64_t)s * 1000000000LL + (int64_t)round(frac);
}

// sorted order (by sensor pixel, stable in time) -> pair flag per sorted position
__global__ void k_pair_flags(const uint32_t* __restrict__ skey, int64_t N, int32_t* __restrict__ flag) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  flag[j] = (j > 0 && skey[j] == skey[j - 1]) ? 1 : 0;
}

__global__ void k_pair_links(const uint32_t* __restrict__ sid, const int32_t* __restrict__ flag,
                             const int32_t* __restrict__ rank, int64_t N, int32_t* __restrict__ prev,
                             uint32_t* __restrict__ refrank, int64_t* __restrict__ total) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint32_t ev = sid[j];
  if (flag[j]) {
    prev[ev] = (int32_t)sid[j - 1];
    refrank[ev] = (uint32_t)rank[j];
  } else {
    prev[ev] = -1;
    refrank[ev] = 0xFFFFFFFFu;
  }
  if (j == N - 1) *total = (int64_t)rank[j] + flag[j];
}

// ---- several GPUs: every rank pairs only its own time slice of the events; a pair whose previous event lies in an
// earlier slice is closed through the per-sensor-pixel "last event" tables of the ranks (halo)
__global__ void k_pix_tables(const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sid, int64_t N, int64_t ev_off,
                             int32_t* __restrict__ last, int32_t* __restrict__ firstpos, int32_t* __restrict__ lastpos) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint32_t s = skey[j];
  if (j == 0 || skey[j - 1] != s) firstpos[s] = (int32_t)j;
  if (j == N - 1 || skey[j + 1] != s) { lastpos[s] = (int32_t)j; last[s] = (int32_t)(ev_off + sid[j]); }
}
// last event of the pixel in the nearest earlier slice that has one
__global__ void k_halo(int S, int rank, const int32_t* __restrict__ last_all, int32_t* __restrict__ halo) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  int32_t v = -1;
  for (int r = rank - 1; r >= 0 && v < 0; r--) v = last_all[(size_t)r * (S + 1) + s];
  halo[s] = v;
}
__global__ void k_pair_flags_halo(const uint32_t* __restrict__ skey, int64_t N, const int32_t* __restrict__ halo,
                                  int32_t* __restrict__ flag) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const bool first = j == 0 || skey[j] != skey[j - 1];
  flag[j] = (!first || halo[skey[j]] >= 0) ? 1 : 0;
}
// pairs of every sensor pixel in this slice
__global__ void k_pix_counts(int S, const int32_t* __restrict__ firstpos, const int32_t* __restrict__ lastpos,
                             const int32_t* __restrict__ flag, const int32_t* __restrict__ rank, int32_t* __restrict__ cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int f = firstpos[s], l = lastpos[s];
  cnt[s] = f >= 0 ? rank[l] + flag[l] - rank[f] : 0;
}
// tot[s] = pairs of pixel s over all slices, lower[s] = those of the earlier slices
__global__ void k_pix_totals(int S, int rank, int W, const int32_t* __restrict__ cnt_all, int32_t* __restrict__ tot,
                             int32_t* __restrict__ lower) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > S) return;
  int t = 0, lo = 0;
  if (s < S)
    for (int r = 0; r < W; r++) { const int c = cnt_all[(size_t)r * (S + 1) + s]; t += c; if (r < rank) lo += c; }
  tot[s] = t;  // tot[S] = 0: the scan's last entry becomes the window's pair count
  if (s < S) lower[s] = lo;
}
__global__ void k_pair_links_halo(const uint32_t* __restrict__ skey, const uint32_t* __restrict__ sid,
                                  const int32_t* __restrict__ flag, const int32_t* __restrict__ rank,
                                  const int32_t* __restrict__ firstpos, const int32_t* __restrict__ halo,
                                  const int32_t* __restrict__ pixbase, const int32_t* __restrict__ lower, int64_t N,
                                  int64_t ev_off, int32_t* __restrict__ prev, uint32_t* __restrict__ refrank,
                                  int64_t* __restrict__ total) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const uint32_t ev = sid[j], s = skey[j];
  if (flag[j]) {
    const bool first = j == 0 || skey[j - 1] != s;
    prev[ev] = first ? halo[s] : (int32_t)(ev_off + sid[j - 1]);
    refrank[ev] = (uint32_t)(pixbase[s] + lower[s] + (rank[j] - rank[firstpos[s]]));
  } else {
    prev[ev] = -1;
    refrank[ev] = 0xFFFFFFFFu;
  }
  if (j == N - 1) *total = (int64_t)rank[j] + flag[j];
}

// knot index and normalised time of every batch mid-time: basalt So3Spline::evaluate,
// reference thirdparty/basalt-headers/include/basalt/spline/so3_spline.h:219-230
__global__ void k_batch_su(const int64_t* __restrict__ tmid, int64_t B, int64_t t0, int64_t dt, int n,
                           int32_t* __restrict__ bs, double* __restrict__ bu, int32_t* __restrict__ flags) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t st = tmid[b] - t0;
  int64_t s = st / dt;
  const double u = (double)(st % dt) / (double)dt;
  if (st < 0 || s < 0 || s + 2 > (int64_t)n) {
    atomicOr(flags, 2);
    s = 0;
  }
  bs[b] = (int32_t)s;
  bu[b] = u;
}

// per event (time order): is it the current event of a pair
__global__ void k_meas_flags(const int32_t* __restrict__ prev, int64_t N, int32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  flag[i] = prev[i] >= 0 ? 1 : 0;
}

// compacted (control-pose-pair key, event) list in time order
__global__ void k_meas_compact(const int32_t* __restrict__ prev, const int32_t* __restrict__ bs,
                               const int32_t* __restrict__ pos, int64_t N, int n, int64_t ev_off,
                               uint32_t* __restrict__ key, uint32_t* __restrict__ ev) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int32_t p = prev[i];  // global id
  if (p < 0) return;
  const uint32_t cc = (uint32_t)bs[(ev_off + i) / kBatch];
  const uint32_t cp = (uint32_t)bs[p / kBatch];
  const int32_t o = pos[i];
  key[o] = cc * (uint32_t)n + cp;  // n <= 65535 (checked by emba_set_state): fits 32 bits
  ev[o] = (uint32_t)i;
}

__global__ void k_head_flags(const uint32_t* __restrict__ key, int64_t M, int32_t* __restrict__ flag) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  flag[j] = (j == 0 || key[j] != key[j - 1]) ? 1 : 0;
}

__global__ void k_head_scatter(const uint32_t* __restrict__ key, const int32_t* __restrict__ flag,
                               const int32_t* __restrict__ gidx, int64_t M, int32_t* __restrict__ gstart,
                               uint32_t* __restrict__ gkey, int64_t* __restrict__ total) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  if (flag[j]) {
    gstart[gidx[j]] = (int32_t)j;
    gkey[gidx[j]] = key[j];
  }
  if (j == M - 1) *total = (int64_t)gidx[j] + flag[j];
}

__global__ void k_build_recs(const uint32_t* __restrict__ sev, int64_t m_lo, int64_t Mloc,
                             const uint32_t* __restrict__ spix, const uint8_t* __restrict__ pol,
                             const int32_t* __restrict__ prev, const uint32_t* __restrict__ refrank,
                             const double* __restrict__ lut, int64_t ev_off, MeasRec* __restrict__ rec,
                             uint32_t* __restrict__ refpos) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Mloc) return;
  const uint32_t ev = sev[m_lo + j];  // local id of the current event; prev[] holds global ids
  MeasRec r;
  const size_t sp = spix[ev];
  // unit bearing: rotations keep the norm, and both the projection (atan2 of a ratio, asin of y / norm) and its
  // Jacobian are homogeneous of degree 0 in the bearing, so the per-measurement norm and division
  // (equirectangular_camera.h:22-24) are done once here instead of twice per evaluation
  const double lx = lut[3 * sp], ly = lut[3 * sp + 1], lz = lut[3 * sp + 2];
  const double inv = 1.0 / sqrt(lx * lx + ly * ly + lz * lz);
  r.bx = lx * inv; r.by = ly * inv; r.bz = lz * inv;
  r.bc_pol = (uint32_t)((ev_off + ev) / kBatch) | ((uint32_t)(pol[ev] ? 1u : 0u) << 31);
  r.bp = (uint32_t)prev[ev] / kBatch;
  rec[j] = r;
  refpos[j] = refrank[ev];
}

static int bits_for(uint64_t maxval) {
  int b = 1;
  while (b < 32 && (maxval >> b)) b++;
  return b;
}

int rebuild_static(Handle* h);
void comm_destroy(Handle* h);
struct PoissonPlan;
void poisson_plan_destroy(PoissonPlan* p);

int comm_allgather_i32(Handle* h, const int32_t* send, int32_t* recv, int64_t count);

// The per-window pre-pass on device-resident coordinates. d_x / d_y / d_pol_src: this rank's Nloc events (device);
// d_tpair: the first and last timestamp of every batch of the WINDOW (device, 2 B entries).
static int prepass_device(Handle* h, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_pol_src,
                          const int64_t* d_tpair) {
  const int64_t Nu = h->Nloc, B = h->B;
  const int64_t Nu1 = std::max<int64_t>(Nu, 1);
  const int T = 256, G = ceil_div64(Nu1, T);
  const int S = h->Ws * h->Hs, W = h->world;
  // scratch: sort buffers (4 x u32), pair flags + ranks, radix / scan scratch, per-sensor-pixel tables (several GPUs)
  uint32_t* k0 = h->ar_tmp.take<uint32_t>(Nu1);
  uint32_t* v0 = h->ar_tmp.take<uint32_t>(Nu1);
  uint32_t* k1 = h->ar_tmp.take<uint32_t>(Nu1);
  uint32_t* v1 = h->ar_tmp.take<uint32_t>(Nu1);
  int32_t* d_flag = h->ar_tmp.take<int32_t>(Nu1);
  int32_t* d_rank = h->ar_tmp.take<int32_t>(Nu1);
  void* scr = h->ar_tmp.take<char>((int64_t)std::max(radix_scratch_bytes(Nu1), scan_scratch_bytes(std::max<int64_t>(Nu1, S + 2))));
  int64_t* d_total = h->ar_tmp.take<int64_t>(4);
  int32_t* tab = W > 1 ? h->ar_tmp.take<int32_t>((int64_t)(7 + 2 * W) * (S + 1)) : nullptr;
  if (!k0 || !v0 || !k1 || !v1 || !d_flag || !d_rank || !scr || !d_total || (W > 1 && !tab)) { h->err = "pre-pass scratch arena too small"; return EMBA_E_CUDA; }
  EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
  EMBA_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t) * 4, h->stream));
  if (Nu && d_pol_src != h->d_pol) EMBA_CUDA(cudaMemcpyAsync(h->d_pol, d_pol_src, Nu, cudaMemcpyDeviceToDevice, h->stream));
  if (B) {
    k_batch_mid<<<ceil_div64(B, T), T, 0, h->stream>>>(d_tpair, B, h->d_tmid);
    EMBA_LAUNCH_CHECK();
  }
  const uint32_t *ks = k0, *vs = v0;
  if (Nu) {
    k_spix<<<G, T, 0, h->stream>>>(d_x, d_y, h->Ws, h->Hs, Nu, h->d_spix_ev, k0, h->d_flags);
    EMBA_LAUNCH_CHECK();
    int which = 0;
    EMBA_TRY(radix_sort_pairs(h, h->stream, k0, v0, k1, v1, Nu, bits_for((uint64_t)S - 1), scr, true, true, &which));
    ks = which ? k1 : k0;
    vs = which ? v1 : v0;
  }
  if (W == 1) {
    if (Nu) {
      k_pair_flags<<<G, T, 0, h->stream>>>(ks, Nu, d_flag);
      EMBA_LAUNCH_CHECK();
      EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, d_flag, d_rank, Nu, scr));
      k_pair_links<<<G, T, 0, h->stream>>>(vs, d_flag, d_rank, Nu, h->d_prev, h->d_refrank, d_total);
      EMBA_LAUNCH_CHECK();
    }
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 8, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  } else {
    // per-sensor-pixel tables: last event / first and last sorted position in my slice; the ranks exchange the
    // "last event" tables (halo) and, once the halo pairs are known, their per-pixel pair counts (reference-order
    // ranks of the pairs, window total)
    const int S1 = S + 1;
    int32_t *last = tab, *firstpos = tab + S1, *lastpos = tab + 2 * S1, *halo = tab + 3 * S1, *cnt = tab + 4 * S1,
            *tot = tab + 5 * S1, *lower = tab + 6 * S1, *last_all = tab + 7 * S1, *cnt_all = last_all + (size_t)W * S1;
    EMBA_CUDA(cudaMemsetAsync(tab, 0xFF, sizeof(int32_t) * 3 * S1, h->stream));  // last, firstpos, lastpos = -1
    if (Nu) {
      k_pix_tables<<<G, T, 0, h->stream>>>(ks, vs, Nu, h->ev_off, last, firstpos, lastpos);
      EMBA_LAUNCH_CHECK();
    }
    EMBA_TRY(comm_allgather_i32(h, last, last_all, S1));
    k_halo<<<ceil_div64(S, T), T, 0, h->stream>>>(S, h->rank, last_all, halo);
    EMBA_LAUNCH_CHECK();
    // (the tables above are strided by S + 1: last_all[r * (S + 1) + s])
    if (Nu) {
      k_pair_flags_halo<<<G, T, 0, h->stream>>>(ks, Nu, halo, d_flag);
      EMBA_LAUNCH_CHECK();
      EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, d_flag, d_rank, Nu, scr));
    }
    k_pix_counts<<<ceil_div64(S, T), T, 0, h->stream>>>(S, firstpos, lastpos, d_flag, d_rank, cnt);
    EMBA_LAUNCH_CHECK();
    EMBA_CUDA(cudaMemsetAsync(cnt + S, 0, sizeof(int32_t), h->stream));
    EMBA_TRY(comm_allgather_i32(h, cnt, cnt_all, S1));
    k_pix_totals<<<ceil_div64(S1, T), T, 0, h->stream>>>(S, h->rank, W, cnt_all, tot, lower);
    EMBA_LAUNCH_CHECK();
    EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, tot, tot, S1, scr));  // tot -> first reference rank of the pixel
    if (Nu) {
      k_pair_links_halo<<<G, T, 0, h->stream>>>(ks, vs, d_flag, d_rank, firstpos, halo, tot, lower, Nu, h->ev_off, h->d_prev,
                                                h->d_refrank, d_total);
      EMBA_LAUNCH_CHECK();
    }
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 8, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 10, tot + S, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  }
  EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 9, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));  // the one synchronisation of the event-level pre-pass
  if (*reinterpret_cast<int32_t*>(h->h_pin + 9) & 1) { h->err = "emba_set_events: event coordinates outside the sensor"; return EMBA_E_ARG; }
  h->Mloc_pairs = h->h_pin[8];
  h->Mc_total = W == 1 ? h->Mloc_pairs : (int64_t)*reinterpret_cast<int32_t*>(h->h_pin + 10);
  return EMBA_OK;
}

// sizes the arenas of a window of N events and lays out the event-level arrays
static int begin_window(Handle* h, int64_t N) {
  h->N = N;
  h->Nuse = (N / kBatch) * kBatch;  // integer division at model.cpp:79 drops the tail batch
  h->B = h->Nuse / kBatch;
  // this rank's slice: whole batches, split by event COUNT (load balance by count, not by time)
  const int64_t b_lo = h->B * h->rank / h->world, b_hi = h->B * (h->rank + 1) / h->world;
  h->ev_off = b_lo * kBatch;
  h->Nloc = (b_hi - b_lo) * kBatch;
  h->t0_ns = -1;  // forces the spline-dependent structures to be rebuilt
  h->st[0].evaluated = h->st[1].evaluated = false;
  h->formed = h->solved = false;
  h->jrec_valid = false;
  h->Mc_total = 0;
  h->Mloc_pairs = 0;
  h->Mc = 0;
  const int64_t Nu = std::max<int64_t>(h->Nloc, 1), B = std::max<int64_t>(h->B, 1);
  const int64_t S1 = (int64_t)h->Ws * h->Hs + 1;
  const size_t ev_bytes = Arena::pad(4 * (size_t)Nu) * 3 + Arena::pad((size_t)Nu) + Arena::pad(8 * (size_t)B) * 3 +
                          Arena::pad(16 * (size_t)B) + Arena::pad(4 * (size_t)B) + 4096;
  EMBA_TRY(arena_reserve(h, h->ar_ev, ev_bytes));
  h->d_spix_ev = h->ar_ev.take<uint32_t>(Nu);
  h->d_prev = h->ar_ev.take<int32_t>(Nu);
  h->d_refrank = h->ar_ev.take<uint32_t>(Nu);
  h->d_pol = h->ar_ev.take<uint8_t>(Nu);
  h->d_tmid = h->ar_ev.take<int64_t>(B);
  h->d_bu = h->ar_ev.take<double>(B);
  h->d_bs = h->ar_ev.take<int32_t>(B);
  if (!h->d_spix_ev || !h->d_prev || !h->d_refrank || !h->d_pol || !h->d_tmid || !h->d_bu || !h->d_bs) {
    h->err = "event arena too small"; return EMBA_E_CUDA;
  }
  // scratch: the larger of the pairing pass (6 x 4 N + sort scratch + staged x, y + sensor-pixel tables) and the
  // static rebuild (below)
  const size_t tmp_bytes = Arena::pad(4 * (size_t)Nu) * 6 + Arena::pad(2 * (size_t)Nu) * 2 + Arena::pad(16 * (size_t)B) +
                           Arena::pad(std::max(radix_scratch_bytes(Nu), scan_scratch_bytes(std::max<int64_t>(Nu, S1 + 1)))) +
                           Arena::pad(4 * (size_t)(7 + 2 * h->world) * (size_t)S1) + 16384;
  EMBA_TRY(arena_reserve(h, h->ar_tmp, tmp_bytes));
  return EMBA_OK;
}

}  // namespace emba

using namespace emba;

extern "C" {

const char* emba_version(void) { return "emba_b200 0.2 sm_100a"; }

int emba_create(const emba_config_t* cfg, emba_handle_t* out) {
  if (!cfg || !out || !cfg->bearing_lut || cfg->sensor_w <= 0 || cfg->sensor_h <= 0 || cfg->pano_w <= 0 ||
      cfg->pano_h <= 0)
    return EMBA_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
    return EMBA_E_CUDA;  // no CUDA device: there is no CPU fallback
  Handle* h = new Handle();
  h->device = cfg->device;
  if (cudaSetDevice(h->device) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  h->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  {
    // the side stream (row placement / segment sort beside the pose-side kernel, A11 gather beside the map-side kernel)
    // has the main stream's priority. Giving it the HIGHER priority (EMBA_SIDE_PRIO=1) was measured on C4: its kernels
    // then take every SM slot the HBM-bound pose-side kernel frees and stretch that kernel from 7.0 to 11.5 ms
    // (form 13.4 -> 16.1 ms).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const bool hi = getenv("EMBA_SIDE_PRIO") && atoi(getenv("EMBA_SIDE_PRIO")) == 1;
    if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi ? prio_hi : prio_lo) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_host, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming);
  if (cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  for (auto& e : h->ev_chunk) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (auto& e : h->ev_x) cudaEventCreate(&e);
  if (cudaMallocHost((void**)&h->h_pin, 1024 * sizeof(int64_t)) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  if (uploader_init(h->up) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  cudaEventCreate(&h->ev_sort0);
  cudaEventCreate(&h->ev_sort1);
  for (auto& e : h->ev) cudaEventCreate(&e);
  h->Ws = cfg->sensor_w; h->Hs = cfg->sensor_h; h->Wp = cfg->pano_w; h->Hp = cfg->pano_h;
  h->P = (int64_t)h->Wp * h->Hp;
  h->C_th = cfg->C_th;
  const size_t lut_bytes = sizeof(double) * 3 * (size_t)h->Ws * h->Hs;
  bool ok = cudaMalloc((void**)&h->d_lut, lut_bytes) == cudaSuccess &&
            cudaMemcpy(h->d_lut, cfg->bearing_lut, lut_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMalloc((void**)&h->d_scal, sizeof(double) * 64) == cudaSuccess &&
            cudaMalloc((void**)&h->d_flags, sizeof(int32_t) * 16) == cudaSuccess &&
            cudaMemset(h->d_flags, 0, sizeof(int32_t) * 16) == cudaSuccess;
  for (int s = 0; s < 2 && ok; s++) {
    StateSlot& st = h->st[s];
    ok = ok && cudaMalloc((void**)&st.Gx, sizeof(double) * h->P) == cudaSuccess &&
         cudaMalloc((void**)&st.Gy, sizeof(double) * h->P) == cudaSuccess &&
         cudaMalloc((void**)&st.G2, sizeof(double2) * h->P) == cudaSuccess &&
         cudaMalloc((void**)&st.H3, sizeof(double4) * h->P) == cudaSuccess &&
         cudaMalloc((void**)&st.hist, sizeof(int32_t) * h->P) == cudaSuccess;
    st.hist_loc = st.hist;
  }
  // per-pixel buffers of the normal equations are sized for the worst case (every pixel active) once, so that
  // forming the equations never allocates
  const size_t P1 = (size_t)h->P + 1;
  ok = ok && cudaMalloc((void**)&h->d_amap, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_pflag, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_paidx, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_len, sizeof(int64_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_apix, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_segoff, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_segend, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_segcnt, sizeof(int32_t) * (P1 + 1)) == cudaSuccess &&
       cudaMalloc((void**)&h->d_longlist, sizeof(int32_t) * (P1 + 1)) == cudaSuccess &&
       cudaMalloc((void**)&h->d_scan_tmp, scan_scratch_bytes((int64_t)P1 + 1) * 2) == cudaSuccess &&
       cudaMalloc((void**)&h->d_gmask, sizeof(unsigned long long) * 2 * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_gmask2, sizeof(unsigned long long) * 2 * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_win64, sizeof(int2) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_winlo, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_winhi, sizeof(int32_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_stripoff, sizeof(int64_t) * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_A22, sizeof(double) * 3 * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_b2, sizeof(double) * 2 * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_C, sizeof(double) * 3 * P1) == cudaSuccess &&
       cudaMalloc((void**)&h->d_x2, sizeof(double) * 2 * P1) == cudaSuccess;
  if (!ok) { cudaGetLastError(); emba_destroy((emba_handle_t)h); return EMBA_E_CUDA; }
  *out = (emba_handle_t)h;
  return EMBA_OK;
}

int emba_destroy(emba_handle_t hh) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_OK;
  cudaSetDevice(h->device);
  if (h->stream3) cudaStreamSynchronize(h->stream3);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_destroy(h);
  if (h->poisson) { poisson_plan_destroy((PoissonPlan*)h->poisson); h->poisson = nullptr; }
  for (int s = 0; s < 2; s++) {
    StateSlot& st = h->st[s];
    if (st.hist_loc && st.hist_loc != st.hist) cudaFree(st.hist_loc);
    cudaFree(st.Gx); cudaFree(st.Gy); cudaFree(st.G2); cudaFree(st.H3); cudaFree(st.hist);
    st = StateSlot();
  }
  void* ptrs[] = {h->d_lut, h->ar_ev.base, h->ar_meas.base, h->ar_tmp.base, h->d_part, h->d_scal, h->d_flags, h->d_amap,
                  h->d_pflag, h->d_paidx, h->d_len, h->d_apix, h->d_segoff, h->d_segend, h->d_segcnt, h->d_longlist,
                  h->d_scan_tmp, h->d_gmask, h->d_gmask2, h->d_win64, h->d_winlo, h->d_winhi, h->d_stripoff, h->d_strip,
                  h->d_A22, h->d_b2, h->d_A11, h->d_b1, h->d_C, h->d_S, h->d_rhs, h->d_x1, h->d_x2, h->d_Spart, h->d_cg,
                  h->d_ldlt_w, h->d_win2, h->d_win_all, h->d_own_len, h->d_gwinlo, h->d_gwinhi, h->d_gstripoff,
                  h->d_gstrip, h->d_recv, h->d_dst, h->d_peerx};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {h->ev_fork, h->ev_join, h->ev_fork2, h->ev_join2, h->ev_sort0, h->ev_sort1, h->ev_host}) if (e) cudaEventDestroy(e);
  uploader_free(h->up);
  for (auto& e : h->ev_chunk) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_x) if (e) cudaEventDestroy(e);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  if (h->d_glen) cudaFree(h->d_glen);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->stream3) cudaStreamDestroy(h->stream3);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EMBA_OK;
}

const char* emba_last_error(emba_handle_t hh) {
  Handle* h = (Handle*)hh;
  return h ? h->err.c_str() : "null handle";
}

int emba_set_shard(emba_handle_t hh, int32_t rank, int32_t world) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (world < 1 || rank < 0 || rank >= world) { h->err = "bad shard"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  if (rank != h->rank || world != h->world) { h->rank = rank; h->world = world; h->t0_ns = -1; }
  // several ranks: the evaluation counts into a rank-local histogram (it sizes the rank's row segments), the
  // all-reduce writes the global one beside it
  for (int s = 0; s < 2; s++) {
    StateSlot& st = h->st[s];
    if (world > 1 && st.hist_loc == st.hist) {
      st.hist_loc = nullptr;
      EMBA_TRY(dev_alloc(h, &st.hist_loc, h->P));
    }
  }
  return EMBA_OK;
}

int emba_set_strict_range(emba_handle_t hh, int32_t on) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  h->strict_range = on ? 1 : 0;
  return EMBA_OK;
}

int emba_set_events(emba_handle_t hh, int64_t N, const uint16_t* x, const uint16_t* y, const int64_t* t_ns,
                    const uint8_t* pol) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (N < 0 || (N > 0 && (!x || !y || !t_ns || !pol))) { h->err = "emba_set_events: null input"; return EMBA_E_ARG; }
  if (N >= (int64_t)1 << 31) { h->err = "emba_set_events: more than 2^31-1 events per window"; return EMBA_E_ARG; }
  if (h->world > 1 && !h->nccl_comm) { h->err = "emba_set_events: several ranks need a communicator first (emba_comm_init)"; return EMBA_E_NCCL; }
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  EMBA_TRY(begin_window(h, N));
  const int64_t Nl = h->Nloc, B = h->B, off = h->ev_off;
  if (h->Nuse == 0) return EMBA_OK;
  uint16_t* d_x = h->ar_tmp.take<uint16_t>(std::max<int64_t>(Nl, 1));
  uint16_t* d_y = h->ar_tmp.take<uint16_t>(std::max<int64_t>(Nl, 1));
  int64_t* d_tpair = h->ar_tmp.take<int64_t>(2 * B);
  if (!d_x || !d_y || !d_tpair) { h->err = "pre-pass scratch arena too small"; return EMBA_E_CUDA; }
  // a rank uploads only its own slice of the events. Timestamps stay on the host: only the first and the last stamp
  // of every batch of the window are needed (model.cpp:115-119); they are gathered into the pinned stage in chunks.
  EMBA_CUDA(upload_bytes(h->up, h->stream, d_x, x + off, sizeof(uint16_t) * (size_t)Nl));
  EMBA_CUDA(upload_bytes(h->up, h->stream, d_y, y + off, sizeof(uint16_t) * (size_t)Nl));
  EMBA_CUDA(upload_bytes(h->up, h->stream, h->d_pol, pol + off, (size_t)Nl));
  {
    const int64_t per = (int64_t)(kStageBytes / 16);
    for (int64_t b0 = 0; b0 < B; b0 += per) {
      const int64_t nb = std::min(per, B - b0);
      const int bi = h->up.turn;
      h->up.turn ^= 1;
      EMBA_CUDA(uploader_init(h->up));
      EMBA_CUDA(cudaEventSynchronize(h->up.ev[bi]));
      int64_t* st = reinterpret_cast<int64_t*>(h->up.stage[bi]);
      for (int64_t b = 0; b < nb; b++) {
        st[2 * b] = t_ns[(b0 + b) * kBatch];
        st[2 * b + 1] = t_ns[(b0 + b) * kBatch + kBatch - 1];
      }
      EMBA_CUDA(cudaMemcpyAsync(d_tpair + 2 * b0, st, sizeof(int64_t) * 2 * nb, cudaMemcpyHostToDevice, h->stream));
      EMBA_CUDA(cudaEventRecord(h->up.ev[bi], h->stream));
    }
  }
  EMBA_TRY(prepass_device(h, d_x, d_y, h->d_pol, d_tpair));
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  EMBA_CUDA(cudaEventSynchronize(h->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->t_setup_ms[0] = ms;
  return EMBA_OK;
}

int emba_num_pairs(emba_handle_t hh, int64_t* out) {
  Handle* h = (Handle*)hh;
  if (!h || !out) return EMBA_E_ARG;
  *out = h->Mc_total;
  return EMBA_OK;
}

// ---------------------------------------------------------------------------------------------------
// device-resident event sequence (N1)
// ---------------------------------------------------------------------------------------------------
#define EV_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ev->err = std::string(#call) + ": " + cudaGetErrorString(_e);                        \
      return EMBA_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

int emba_events_create(int32_t device, int64_t N, const uint16_t* x, const uint16_t* y, const int64_t* t_ns,
                       const uint8_t* pol, emba_events_t* out) {
  if (!out || N < 0 || (N > 0 && (!x || !y || !t_ns || !pol))) return EMBA_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return EMBA_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return EMBA_E_CUDA;
  EventStore* ev = new EventStore();
  ev->device = device;
  ev->N = N;
  ev->cap = std::max<int64_t>(N, 1);
  bool ok = cudaStreamCreateWithFlags(&ev->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaMallocHost((void**)&ev->h_pin, 64 * sizeof(int64_t)) == cudaSuccess &&
            cudaMalloc((void**)&ev->x, sizeof(uint16_t) * ev->cap) == cudaSuccess &&
            cudaMalloc((void**)&ev->y, sizeof(uint16_t) * ev->cap) == cudaSuccess &&
            cudaMalloc((void**)&ev->t, sizeof(int64_t) * ev->cap) == cudaSuccess &&
            cudaMalloc((void**)&ev->pol, sizeof(uint8_t) * ev->cap) == cudaSuccess;
  ok = ok && upload_bytes(ev->up, ev->stream, ev->x, x, sizeof(uint16_t) * (size_t)N) == cudaSuccess &&
       upload_bytes(ev->up, ev->stream, ev->y, y, sizeof(uint16_t) * (size_t)N) == cudaSuccess &&
       upload_bytes(ev->up, ev->stream, ev->t, t_ns, sizeof(int64_t) * (size_t)N) == cudaSuccess &&
       upload_bytes(ev->up, ev->stream, ev->pol, pol, (size_t)N) == cudaSuccess &&
       cudaStreamSynchronize(ev->stream) == cudaSuccess;
  if (!ok) { cudaGetLastError(); emba_events_destroy((emba_events_t)ev); return EMBA_E_CUDA; }
  *out = (emba_events_t)ev;
  return EMBA_OK;
}

int emba_events_destroy(emba_events_t e) {
  EventStore* ev = (EventStore*)e;
  if (!ev) return EMBA_OK;
  cudaSetDevice(ev->device);
  if (ev->stream) cudaStreamSynchronize(ev->stream);
  cudaFree(ev->x); cudaFree(ev->y); cudaFree(ev->t); cudaFree(ev->pol);
  uploader_free(ev->up);
  if (ev->h_pin) cudaFreeHost(ev->h_pin);
  if (ev->stream) cudaStreamDestroy(ev->stream);
  delete ev;
  return EMBA_OK;
}

int emba_events_count(emba_events_t e, int64_t* out) {
  EventStore* ev = (EventStore*)e;
  if (!ev || !out) return EMBA_E_ARG;
  *out = ev->N;
  return EMBA_OK;
}

int emba_events_download(emba_events_t e, int64_t i0, int64_t i1, uint16_t* x, uint16_t* y, int64_t* t_ns,
                         uint8_t* pol) {
  EventStore* ev = (EventStore*)e;
  if (!ev || i0 < 0 || i1 < i0 || i1 > ev->N) return EMBA_E_ARG;
  EV_CUDA(cudaSetDevice(ev->device));
  const size_t n = (size_t)(i1 - i0);
  if (n == 0) return EMBA_OK;
  if (x) EV_CUDA(cudaMemcpyAsync(x, ev->x + i0, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, ev->stream));
  if (y) EV_CUDA(cudaMemcpyAsync(y, ev->y + i0, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, ev->stream));
  if (t_ns) EV_CUDA(cudaMemcpyAsync(t_ns, ev->t + i0, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ev->stream));
  if (pol) EV_CUDA(cudaMemcpyAsync(pol, ev->pol + i0, n, cudaMemcpyDeviceToHost, ev->stream));
  EV_CUDA(cudaStreamSynchronize(ev->stream));
  return EMBA_OK;
}

}  // extern "C"

namespace emba {

// ---- sequence-level kernels
__global__ void k_time_minmax_sorted(const int64_t* __restrict__ t, int64_t N, unsigned long long* __restrict__ out) {
  // out[0] = min, out[1] = max, out[2] = number of descents t[i] < t[i-1]
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  long long mn = LLONG_MAX, mx = LLONG_MIN;
  unsigned long long desc = 0;
  for (; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const long long v = t[i];
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
    if (i > 0 && v < t[i - 1]) desc++;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_down_sync(0xffffffffu, mn, o), b = __shfl_down_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
    desc += __shfl_down_sync(0xffffffffu, desc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<long long*>(out), mn);
    atomicMax(reinterpret_cast<long long*>(out + 1), mx);
    atomicAdd(out + 2, desc);
  }
}
__global__ void k_time_key(const int64_t* __restrict__ t, const uint32_t* __restrict__ perm, int64_t N, int64_t tmin,
                           int shift, uint32_t* __restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t src = perm ? perm[i] : i;
  key[i] = (uint32_t)(((unsigned long long)(t[src] - tmin)) >> shift);
}
template <typename T>
__global__ void k_gather(const T* __restrict__ in, const uint32_t* __restrict__ perm, int64_t N, T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = in[perm[i]];
}
template <typename T>
__global__ void k_stride_pick(const T* __restrict__ in, int64_t Nout, int rate, T* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Nout) out[j] = in[(j + 1) * rate - 1];  // the rate-th, 2 rate-th ... event (emba.cpp:288-300)
}

// EMBA::getEventSubset (emba.cpp:473-510) on a time-sorted sequence: the two 100-event-stride linear searches stop
// at the first probe whose stamp exceeds the robust bound, which on sorted stamps is a binary search over the probes.
__global__ void k_window_search(const int64_t* __restrict__ t, int64_t N, int64_t lo_ns, int64_t hi_ns,
                                int64_t* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t nprobe = (N + 99) / 100;  // probes at 0, 100, 200, ... < N
  auto first_above = [&](int64_t from, int64_t bound) {
    int64_t a = from, b = nprobe;
    while (a < b) {
      const int64_t m = (a + b) >> 1;
      if (t[m * 100] > bound) b = m; else a = m + 1;
    }
    return a;  // == nprobe: ran off the end
  };
  const int64_t pb = first_above(0, lo_ns);
  const unsigned long long beg = (unsigned long long)pb * 100ull;  // >= N when no probe lies inside
  unsigned long long end;
  const int64_t pe = first_above(pb < nprobe ? pb : nprobe, hi_ns);
  if (pe < nprobe) end = (unsigned long long)pe * 100ull - 100ull;  // `idx_ev_subset_end -= 100` (size_t: may wrap)
  else end = (unsigned long long)(pe > pb ? pe : pb) * 100ull;
  if (end > (unsigned long long)N) end = (unsigned long long)N;     // emba.cpp:503-504
  out[0] = (int64_t)(beg < (unsigned long long)N ? beg : (unsigned long long)N);
  out[1] = (int64_t)end;
}

}  // namespace emba

extern "C" {

int emba_events_sort_by_time(emba_events_t e) {
  EventStore* ev = (EventStore*)e;
  if (!ev) return EMBA_E_ARG;
  EV_CUDA(cudaSetDevice(ev->device));
  const int64_t N = ev->N;
  if (N < 2) return EMBA_OK;
  if (N >= (int64_t)1 << 32) { ev->err = "more than 2^32-1 events"; return EMBA_E_SUPPORT; }
  Handle hh;  // launch counter / error sink of the primitives
  Handle* h = &hh;
  unsigned long long* d_mm = nullptr;
  EV_CUDA(cudaMalloc((void**)&d_mm, 3 * sizeof(unsigned long long)));
  const unsigned long long init[3] = {(unsigned long long)LLONG_MAX, (unsigned long long)LLONG_MIN, 0ull};
  cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, ev->stream);
  k_time_minmax_sorted<<<1184, 256, 0, ev->stream>>>(ev->t, N, d_mm);
  cudaMemcpyAsync(ev->h_pin, d_mm, sizeof(init), cudaMemcpyDeviceToHost, ev->stream);
  cudaError_t ce = cudaStreamSynchronize(ev->stream);
  cudaFree(d_mm);
  if (ce != cudaSuccess) { ev->err = cudaGetErrorString(ce); return EMBA_E_CUDA; }
  const int64_t tmin = ev->h_pin[0], tmax = ev->h_pin[1];
  if (ev->h_pin[2] == 0) return EMBA_OK;  // already sorted (the usual case: bags are recorded in time order)
  // LSD over the 64-bit span in two 32-bit halves: argsort by the low half, then (stable) by the high half
  const unsigned long long span = (unsigned long long)(tmax - tmin);
  const int bits_hi = span >> 32 ? bits_for(span >> 32) : 0;
  const int bits_lo = bits_hi ? 32 : bits_for(span);
  uint32_t *k0 = nullptr, *v0 = nullptr, *k1 = nullptr, *v1 = nullptr;
  void *scr = nullptr, *tmp = nullptr;
  int rc = EMBA_OK;
  do {
    if (cudaMalloc((void**)&k0, 4 * N) != cudaSuccess || cudaMalloc((void**)&v0, 4 * N) != cudaSuccess ||
        cudaMalloc((void**)&k1, 4 * N) != cudaSuccess || cudaMalloc((void**)&v1, 4 * N) != cudaSuccess ||
        cudaMalloc(&scr, radix_scratch_bytes(N)) != cudaSuccess || cudaMalloc(&tmp, 8 * N) != cudaSuccess) {
      cudaGetLastError(); ev->err = "emba_events_sort_by_time: out of device memory"; rc = EMBA_E_CUDA; break;
    }
    const int T = 256, G = ceil_div64(N, T);
    k_time_key<<<G, T, 0, ev->stream>>>(ev->t, nullptr, N, tmin, 0, k0);
    int which = 0;
    if ((rc = radix_sort_pairs(h, ev->stream, k0, v0, k1, v1, N, bits_lo, scr, true, false, &which))) break;
    uint32_t* perm = which ? v1 : v0;
    if (bits_hi) {
      uint32_t* ka = which ? k0 : k1;  // the buffer pair not holding the permutation
      uint32_t* va = which ? v0 : v1;
      k_time_key<<<G, T, 0, ev->stream>>>(ev->t, perm, N, tmin, 32, which ? k1 : k0);
      // sort (key_hi, perm) pairs: keys in the perm's own buffer pair, ping-pong with the other pair
      uint32_t* kin = which ? k1 : k0;
      int w2 = 0;
      if ((rc = radix_sort_pairs(h, ev->stream, kin, perm, ka, va, N, bits_hi, scr, false, false, &w2))) break;
      perm = w2 ? va : perm;
    }
    // apply the permutation to the four arrays (through tmp)
    k_gather<int64_t><<<G, T, 0, ev->stream>>>(ev->t, perm, N, (int64_t*)tmp);
    cudaMemcpyAsync(ev->t, tmp, 8 * N, cudaMemcpyDeviceToDevice, ev->stream);
    k_gather<uint16_t><<<G, T, 0, ev->stream>>>(ev->x, perm, N, (uint16_t*)tmp);
    cudaMemcpyAsync(ev->x, tmp, 2 * N, cudaMemcpyDeviceToDevice, ev->stream);
    k_gather<uint16_t><<<G, T, 0, ev->stream>>>(ev->y, perm, N, (uint16_t*)tmp);
    cudaMemcpyAsync(ev->y, tmp, 2 * N, cudaMemcpyDeviceToDevice, ev->stream);
    k_gather<uint8_t><<<G, T, 0, ev->stream>>>(ev->pol, perm, N, (uint8_t*)tmp);
    cudaMemcpyAsync(ev->pol, tmp, N, cudaMemcpyDeviceToDevice, ev->stream);
    ce = cudaStreamSynchronize(ev->stream);
    if (ce != cudaSuccess) { ev->err = cudaGetErrorString(ce); rc = EMBA_E_CUDA; }
  } while (0);
  if (rc != EMBA_OK && ev->err.empty()) ev->err = hh.err;
  cudaFree(k0); cudaFree(v0); cudaFree(k1); cudaFree(v1); cudaFree(scr); cudaFree(tmp);
  return rc;
}

int emba_events_subsample(emba_events_t e, int32_t rate) {
  EventStore* ev = (EventStore*)e;
  if (!ev) return EMBA_E_ARG;
  if (rate < 2) return EMBA_OK;  // emba.cpp:282: only rates >= 2 sample
  EV_CUDA(cudaSetDevice(ev->device));
  const int64_t Nout = ev->N / rate;
  if (Nout > 0) {
    void* tmp = nullptr;
    EV_CUDA(cudaMalloc(&tmp, 8 * (size_t)Nout));
    const int T = 256, G = ceil_div64(Nout, T);
    k_stride_pick<int64_t><<<G, T, 0, ev->stream>>>(ev->t, Nout, rate, (int64_t*)tmp);
    cudaMemcpyAsync(ev->t, tmp, 8 * Nout, cudaMemcpyDeviceToDevice, ev->stream);
    k_stride_pick<uint16_t><<<G, T, 0, ev->stream>>>(ev->x, Nout, rate, (uint16_t*)tmp);
    cudaMemcpyAsync(ev->x, tmp, 2 * Nout, cudaMemcpyDeviceToDevice, ev->stream);
    k_stride_pick<uint16_t><<<G, T, 0, ev->stream>>>(ev->y, Nout, rate, (uint16_t*)tmp);
    cudaMemcpyAsync(ev->y, tmp, 2 * Nout, cudaMemcpyDeviceToDevice, ev->stream);
    k_stride_pick<uint8_t><<<G, T, 0, ev->stream>>>(ev->pol, Nout, rate, (uint8_t*)tmp);
    cudaMemcpyAsync(ev->pol, tmp, Nout, cudaMemcpyDeviceToDevice, ev->stream);
    cudaError_t ce = cudaStreamSynchronize(ev->stream);
    cudaFree(tmp);
    if (ce != cudaSuccess) { ev->err = cudaGetErrorString(ce); return EMBA_E_CUDA; }
  }
  ev->N = Nout;
  return EMBA_OK;
}

int emba_events_window(emba_events_t e, int64_t t_beg_ns, int64_t t_end_ns, int64_t* idx_beg, int64_t* idx_end) {
  EventStore* ev = (EventStore*)e;
  if (!ev || !idx_beg || !idx_end) return EMBA_E_ARG;
  EV_CUDA(cudaSetDevice(ev->device));
  if (ev->N == 0) { *idx_beg = *idx_end = 0; return EMBA_OK; }
  int64_t* d_out = nullptr;
  EV_CUDA(cudaMalloc((void**)&d_out, 2 * sizeof(int64_t)));
  // t_epsilon = ros::Duration(1e-3) = 1 000 000 ns exactly (emba.cpp:476-478)
  k_window_search<<<1, 32, 0, ev->stream>>>(ev->t, ev->N, t_beg_ns + 1000000, t_end_ns - 1000000, d_out);
  cudaMemcpyAsync(ev->h_pin, d_out, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ev->stream);
  cudaError_t ce = cudaStreamSynchronize(ev->stream);
  cudaFree(d_out);
  if (ce != cudaSuccess) { ev->err = cudaGetErrorString(ce); return EMBA_E_CUDA; }
  *idx_beg = ev->h_pin[0];
  *idx_end = ev->h_pin[1];
  return EMBA_OK;
}

int emba_set_events_dev(emba_handle_t hh, emba_events_t e, int64_t i0, int64_t i1) {
  Handle* h = (Handle*)hh;
  EventStore* ev = (EventStore*)e;
  if (!h) return EMBA_E_ARG;
  if (!ev || i0 < 0 || i1 < i0 || i1 > ev->N) { h->err = "emba_set_events_dev: bad range"; return EMBA_E_ARG; }
  if (ev->device != h->device) { h->err = "emba_set_events_dev: the sequence lives on another device"; return EMBA_E_ARG; }
  const int64_t N = i1 - i0;
  if (N >= (int64_t)1 << 31) { h->err = "emba_set_events: more than 2^31-1 events per window"; return EMBA_E_ARG; }
  if (h->world > 1 && !h->nccl_comm) { h->err = "emba_set_events_dev: several ranks need a communicator first (emba_comm_init)"; return EMBA_E_NCCL; }
  EMBA_CUDA(cudaSetDevice(h->device));
  EMBA_CUDA(cudaStreamSynchronize(ev->stream));
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  EMBA_TRY(begin_window(h, N));
  const int64_t B = h->B;
  if (h->Nuse == 0) return EMBA_OK;
  int64_t* d_tpair = h->ar_tmp.take<int64_t>(2 * B);
  if (!d_tpair) { h->err = "pre-pass scratch arena too small"; return EMBA_E_CUDA; }
  k_batch_tpair<<<ceil_div64(B, 256), 256, 0, h->stream>>>(ev->t + i0, B, d_tpair);
  EMBA_LAUNCH_CHECK();
  const int64_t o = i0 + h->ev_off;
  EMBA_TRY(prepass_device(h, ev->x + o, ev->y + o, ev->pol + o, d_tpair));
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  EMBA_CUDA(cudaEventSynchronize(h->ev[1]));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->t_setup_ms[0] = ms;
  return EMBA_OK;
}

}  // extern "C"

namespace emba {

// Rebuilds everything that depends on the spline time base (t0, dt, n) or on the shard: batch (s, u), the
// canonical measurement order with its records, groups and work items, and the per-measurement buffers.
int rebuild_static(Handle* h) {
  const int64_t Nu = h->Nloc, B = h->B;
  const int n = h->n;
  h->Mc = 0; h->n_items = 0; h->n_groups = 0; h->dmax = 0;
  h->h_items.clear();
  h->st[0].evaluated = h->st[1].evaluated = false;
  h->formed = h->solved = false;
  h->jrec_valid = false;
  EMBA_CUDA(cudaEventRecord(h->ev[0], h->stream));
  EMBA_TRY(dev_reserve(h, &h->d_A11, &h->A11_cap, (int64_t)9 * n * n));
  EMBA_TRY(dev_reserve(h, &h->d_b1, &h->b1_cap, (int64_t)3 * n));
  EMBA_TRY(dev_reserve(h, &h->d_x1, &h->x1_cap, (int64_t)3 * n));
  EMBA_TRY(dev_reserve(h, &h->d_S, &h->S_cap, (int64_t)(3 * n + 1) * (3 * n + 1)));  // bordered with the rhs row
  EMBA_TRY(dev_reserve(h, &h->d_rhs, &h->rhs_cap, (int64_t)3 * n));
  const int T = 256;
  const int64_t Mt = h->Mloc_pairs;  // pairs of this rank's slice: every one of them is this rank's measurement
  // scratch of the rebuild (the scratch arena was sized for the pairing pass, which is larger: Mt <= Nu)
  h->ar_tmp.reset();
  int32_t* d_flag = h->ar_tmp.take<int32_t>(std::max(Nu, Mt));
  int32_t* d_pos = h->ar_tmp.take<int32_t>(std::max(Nu, Mt));
  uint32_t* d_key = h->ar_tmp.take<uint32_t>(Mt);
  uint32_t* d_ev = h->ar_tmp.take<uint32_t>(Mt);
  uint32_t* d_key2 = h->ar_tmp.take<uint32_t>(Mt);
  uint32_t* d_ev2 = h->ar_tmp.take<uint32_t>(Mt);
  void* scr = h->ar_tmp.take<char>((int64_t)std::max(radix_scratch_bytes(std::max(Nu, Mt)), scan_scratch_bytes(std::max(Nu, Mt))));
  int64_t* d_total = h->ar_tmp.take<int64_t>(4);
  if (Nu > 0 && (!d_flag || !d_pos || !d_key || !d_ev || !d_key2 || !d_ev2 || !scr || !d_total)) {
    h->err = "static rebuild: scratch arena too small"; return EMBA_E_CUDA;
  }
  std::vector<int32_t> gstart;
  std::vector<uint32_t> gkey;
  const uint32_t* vs = nullptr;
  if (B > 0) {
    EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
    k_batch_su<<<ceil_div64(B, T), T, 0, h->stream>>>(h->d_tmid, B, h->t0_ns, h->dt_ns, n, h->d_bs, h->d_bu, h->d_flags);
    EMBA_LAUNCH_CHECK();
  }
  if (B > 0 && Mt > 0) {
    k_meas_flags<<<ceil_div64(Nu, T), T, 0, h->stream>>>(h->d_prev, Nu, d_flag);
    EMBA_LAUNCH_CHECK();
    EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, d_flag, d_pos, Nu, scr));
    k_meas_compact<<<ceil_div64(Nu, T), T, 0, h->stream>>>(h->d_prev, h->d_bs, d_pos, Nu, n, h->ev_off, d_key, d_ev);
    EMBA_LAUNCH_CHECK();
    int which = 0;
    EMBA_TRY(radix_sort_pairs(h, h->stream, d_key, d_ev, d_key2, d_ev2, Mt, bits_for((uint64_t)n * n - 1), scr, false, true, &which));
    const uint32_t* ks = which ? d_key2 : d_key;
    vs = which ? d_ev2 : d_ev;
    // group heads
    k_head_flags<<<ceil_div64(Mt, T), T, 0, h->stream>>>(ks, Mt, d_flag);
    EMBA_LAUNCH_CHECK();
    EMBA_TRY(scan_exclusive<int32_t>(h, h->stream, d_flag, d_pos, Mt, scr));
    // group starts and keys go to the buffers of the sort's other half (free now)
    int32_t* d_gstart = reinterpret_cast<int32_t*>(which ? d_key : d_key2);
    uint32_t* d_gkey = which ? d_ev : d_ev2;
    k_head_scatter<<<ceil_div64(Mt, T), T, 0, h->stream>>>(ks, d_flag, d_pos, Mt, d_gstart, d_gkey, d_total);
    EMBA_LAUNCH_CHECK();
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 8, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaMemcpyAsync(h->h_pin + 9, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
    if (*reinterpret_cast<int32_t*>(h->h_pin + 9) & 2) {
      h->err = "batch mid-time outside the spline support (the reference aborts here: so3_spline.h:221-230)";
      return EMBA_E_SUPPORT;
    }
    const int64_t G = h->h_pin[8];
    gstart.resize(G); gkey.resize(G);
    EMBA_CUDA(cudaMemcpyAsync(gstart.data(), d_gstart, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaMemcpyAsync(gkey.data(), d_gkey, sizeof(uint32_t) * G, cudaMemcpyDeviceToHost, h->stream));
    EMBA_CUDA(cudaStreamSynchronize(h->stream));
  }
  // host: split groups into work items, pick this rank's contiguous slice (time sharding: groups are ordered by
  // cp_c, i.e. by time), build the lookup tables
  const int G = (int)gstart.size();
  std::vector<WorkItem> all;
  static const int64_t item_max = getenv("EMBA_ITEM_MAX") ? atoll(getenv("EMBA_ITEM_MAX")) : kItemMax;
  for (int g = 0; g < G; g++) {
    const int64_t s0 = gstart[g], s1 = (g + 1 < G) ? gstart[g + 1] : Mt;
    const int cc = (int)(gkey[g] / (uint32_t)n), cp = (int)(gkey[g] % (uint32_t)n);
    h->dmax = std::max(h->dmax, cc - cp);
    for (int64_t s = s0; s < s1; s += item_max) {
      WorkItem w;
      w.cp_c = cc; w.cp_p = cp; w.start = (int32_t)s; w.count = (int32_t)std::min<int64_t>(item_max, s1 - s); w.group = g;
      all.push_back(w);
    }
  }
  // (with several GPUs the time sharding already happened at the event level: all of the slice's pairs are mine)
  const int64_t m_lo = 0;
  h->h_items = all;
  h->Mc = Mt;
  // renumber groups locally (dense ids in order of appearance)
  std::vector<int32_t> item0;
  int lastg = -1, ng = 0;
  for (size_t i = 0; i < h->h_items.size(); i++) {
    WorkItem& w = h->h_items[i];
    w.start -= (int32_t)m_lo;
    if (w.group != lastg) { lastg = w.group; item0.push_back((int32_t)i); ng++; }
    w.group = ng - 1;
  }
  item0.push_back((int32_t)h->h_items.size());
  h->n_items = (int)h->h_items.size();
  h->n_groups = ng;
  std::vector<int32_t> gid((size_t)n * (h->dmax + 1), -1);
  for (const WorkItem& w : h->h_items) gid[(size_t)w.cp_c * (h->dmax + 1) + (w.cp_c - w.cp_p)] = w.group;
  // ---- measurement-level arena
  const int64_t Mc = h->Mc;
  const size_t Mp = (size_t)std::max<int64_t>(Mc, 1);
  size_t mb = 0;
  mb += Arena::pad(sizeof(MeasRec) * Mp) + Arena::pad(4 * Mp);                     // rec, refpos
  mb += 2 * (Arena::pad(16 * Mp) + Arena::pad(8 * Mp) + 2 * Arena::pad(4 * Mp));   // dp, e, pix, slot  x 2 states
  const size_t sval_len = Mp + 8 * std::min<size_t>((size_t)h->P, Mp) + 8192;      // segments padded to 8 ids + tail pad
  mb += Arena::pad(8 * (size_t)kRecDoubles * Mp) + Arena::pad(4 * sval_len);       // Jacobian rows, sorted row ids
  mb += 2 * (Arena::pad(32 * (size_t)n) + Arena::pad(8 * (size_t)kKnotStride * n) + 2 * Arena::pad(32 * (size_t)std::max<int64_t>(B, 1)));
  mb += Arena::pad(sizeof(WorkItem) * (size_t)std::max(1, h->n_items)) + Arena::pad(4 * gid.size()) +
        Arena::pad(4 * item0.size()) + Arena::pad(8 * (size_t)kAccN * std::max(1, h->n_items)) +
        Arena::pad(8 * (size_t)kAccN * std::max(1, h->n_groups)) + 8192;
  // the arena may move: the static rebuild must not lose the canonical event list, which lives in the scratch arena
  EMBA_TRY(arena_reserve(h, h->ar_meas, mb));
  Arena& A = h->ar_meas;
  h->d_rec = A.take<MeasRec>(Mc);
  h->d_refpos = A.take<uint32_t>(Mc);
  for (int s = 0; s < 2; s++) {
    StateSlot& st = h->st[s];
    st.dp = A.take<double2>(Mc); st.e = A.take<double>(Mc); st.pix = A.take<int32_t>(Mc); st.slot = A.take<int32_t>(Mc);
    st.quat = A.take<double>((int64_t)n * 4);
    st.Ktab = A.take<double>((int64_t)n * kKnotStride);
    st.RotTab = A.take<double4>(B);
    st.JacTab = A.take<double4>(B);
  }
  h->d_jrec = A.take<double>(Mc * kRecDoubles);
  h->jrec_cap = Mc * kRecDoubles;
  h->d_sval = A.take<uint32_t>((int64_t)sval_len);
  h->d_items = A.take<WorkItem>(h->n_items);
  h->d_gid = A.take<int32_t>((int64_t)gid.size());
  h->d_group_item0 = A.take<int32_t>((int64_t)item0.size());
  h->d_acc_part = A.take<double>((int64_t)h->n_items * kAccN);
  h->d_gsum = A.take<double>((int64_t)h->n_groups * kAccN);
  if (!h->d_gsum || !h->d_acc_part || !h->d_group_item0 || !h->d_gid || !h->d_items || !h->d_sval || !h->d_jrec) {
    h->err = "measurement arena too small"; return EMBA_E_CUDA;
  }
  if (h->n_items) EMBA_CUDA(cudaMemcpyAsync(h->d_items, h->h_items.data(), sizeof(WorkItem) * h->n_items, cudaMemcpyHostToDevice, h->stream));
  if (!gid.empty()) EMBA_CUDA(cudaMemcpyAsync(h->d_gid, gid.data(), sizeof(int32_t) * gid.size(), cudaMemcpyHostToDevice, h->stream));
  EMBA_CUDA(cudaMemcpyAsync(h->d_group_item0, item0.data(), sizeof(int32_t) * item0.size(), cudaMemcpyHostToDevice, h->stream));
  if (Mc) {
    k_build_recs<<<ceil_div64(Mc, T), T, 0, h->stream>>>(vs, m_lo, Mc, h->d_spix_ev, h->d_pol, h->d_prev,
                                                        h->d_refrank, h->d_lut, h->ev_off, h->d_rec, h->d_refpos);
    EMBA_LAUNCH_CHECK();
  }
  EMBA_CUDA(cudaEventRecord(h->ev[1], h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));  // the host vectors above go out of scope
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->t_setup_ms[1] = ms;
  return EMBA_OK;
}

}  // namespace emba

extern "C" int emba_last_comm_ms(emba_handle_t hh, double* out8) {
  Handle* h = (Handle*)hh;
  if (!h || !out8) return EMBA_E_ARG;
  for (int i = 0; i < 8; i++) out8[i] = h->t_comm_ms[i];
  out8[7] = h->peer_now ? 1.0 : 0.0;
  return EMBA_OK;
}

extern "C" int emba_last_setup_ms(emba_handle_t hh, double* out2) {
  Handle* h = (Handle*)hh;
  if (!h || !out2) return EMBA_E_ARG;
  out2[0] = h->t_setup_ms[0];
  out2[1] = h->t_setup_ms[1];
  return EMBA_OK;
}
