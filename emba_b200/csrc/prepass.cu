// Static structure of a time window: handle lifetime, event upload, batch mid-times, the per-sensor-pixel
// pair links that replace the reference's EventMap (include/emba/event_map.h:22-113), and -- once the spline
// time base is known -- the canonical measurement order (sorted by the control-pose pair a measurement
// touches) with its work items. Everything here runs once per window / per spline base, never per LM iteration.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cub/cub.cuh>

#include "emba_internal.cuh"

namespace emba {

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
__global__ void k_spix(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, int Ws, int Hs, int64_t N,
                       uint32_t* __restrict__ spix, uint32_t* __restrict__ ids, int32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t xi = x[i], yi = y[i];
  if (xi >= (uint32_t)Ws || yi >= (uint32_t)Hs) {
    atomicOr(flags, 1);
    spix[i] = 0;
  } else {
    spix[i] = yi * (uint32_t)Ws + xi;
  }
  ids[i] = (uint32_t)i;
}

// sorted order (by sensor pixel, stable in time) -> pair flag per sorted position
__global__ void k_pair_flags(const uint32_t* __restrict__ skey, int64_t N, int32_t* __restrict__ flag) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  flag[j] = (j > 0 && skey[j] == skey[j - 1]) ? 1 : 0;
}

__global__ void k_pair_links(const uint32_t* __restrict__ sid, const int32_t* __restrict__ flag,
                             const int32_t* __restrict__ rank, int64_t N, int32_t* __restrict__ prev,
                             uint32_t* __restrict__ refrank) {
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

// per event (time order): is it the current event of a pair, and which control-pose pair does it touch
__global__ void k_meas_keys(const int32_t* __restrict__ prev, const int32_t* __restrict__ bs, int64_t N, int n,
                            int32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  flag[i] = prev[i] >= 0 ? 1 : 0;
}

__global__ void k_meas_compact(const int32_t* __restrict__ prev, const int32_t* __restrict__ bs,
                               const int32_t* __restrict__ pos, int64_t N, int n, uint32_t* __restrict__ key,
                               uint32_t* __restrict__ ev) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int32_t p = prev[i];
  if (p < 0) return;
  const uint32_t cc = (uint32_t)bs[i / kBatch];
  const uint32_t cp = (uint32_t)bs[p / kBatch];
  const int32_t o = pos[i];
  key[o] = cc * (uint32_t)n + cp;
  ev[o] = (uint32_t)i;
}

__global__ void k_head_flags(const uint32_t* __restrict__ key, int64_t M, int32_t* __restrict__ flag) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  flag[j] = (j == 0 || key[j] != key[j - 1]) ? 1 : 0;
}

__global__ void k_head_scatter(const uint32_t* __restrict__ key, const int32_t* __restrict__ flag,
                               const int32_t* __restrict__ gidx, int64_t M, int32_t* __restrict__ gstart,
                               uint32_t* __restrict__ gkey) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  if (flag[j]) {
    gstart[gidx[j]] = (int32_t)j;
    gkey[gidx[j]] = key[j];
  }
}

__global__ void k_build_recs(const uint32_t* __restrict__ sev, int64_t m_lo, int64_t Mloc,
                             const uint32_t* __restrict__ spix, const uint8_t* __restrict__ pol,
                             const int32_t* __restrict__ prev, const uint32_t* __restrict__ refrank,
                             const double* __restrict__ lut, MeasRec* __restrict__ rec,
                             uint32_t* __restrict__ refpos) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Mloc) return;
  const uint32_t ev = sev[m_lo + j];
  MeasRec r;
  const size_t sp = spix[ev];
  // unit bearing: rotations keep the norm, and both the projection (atan2 of a ratio, asin of y / norm) and its
  // Jacobian are homogeneous of degree 0 in the bearing, so the per-measurement norm and division
  // (equirectangular_camera.h:22-24) are done once here instead of twice per evaluation
  const double lx = lut[3 * sp], ly = lut[3 * sp + 1], lz = lut[3 * sp + 2];
  const double inv = 1.0 / sqrt(lx * lx + ly * ly + lz * lz);
  r.bx = lx * inv; r.by = ly * inv; r.bz = lz * inv;
  r.bc_pol = (ev / kBatch) | ((uint32_t)(pol[ev] ? 1u : 0u) << 31);
  r.bp = (uint32_t)prev[ev] / kBatch;
  rec[j] = r;
  refpos[j] = refrank[ev];
}

static int bits_for(uint64_t maxval) {
  int b = 1;
  while (b < 32 && (maxval >> b)) b++;
  return b;
}

// stable radix sort of (key, value) pairs; results in the arrays returned through kout/vout
static int sort_pairs(Handle* h, uint32_t* kin, uint32_t* vin, uint32_t* kalt, uint32_t* valt, int64_t count,
                      int end_bit, uint32_t** kout, uint32_t** vout) {
  cub::DoubleBuffer<uint32_t> dk(kin, kalt), dv(vin, valt);
  size_t tmp = 0;
  EMBA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)count, 0, end_bit, h->stream));
  void* d_tmp = nullptr;
  EMBA_CUDA(cudaMalloc(&d_tmp, tmp ? tmp : 1));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp, dk, dv, (int)count, 0, end_bit, h->stream);
  h->launches += 1 + (end_bit + 7) / 8 * 2;
  cudaStreamSynchronize(h->stream);
  cudaFree(d_tmp);
  if (e != cudaSuccess) {
    h->err = std::string("cub sort: ") + cudaGetErrorString(e);
    return EMBA_E_CUDA;
  }
  *kout = dk.Current();
  *vout = dv.Current();
  return EMBA_OK;
}

static int exclusive_sum(Handle* h, const int32_t* in, int32_t* out, int64_t count) {
  size_t tmp = 0;
  EMBA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)count, h->stream));
  void* d_tmp = nullptr;
  EMBA_CUDA(cudaMalloc(&d_tmp, tmp ? tmp : 1));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp, in, out, (int)count, h->stream);
  h->launches += 2;
  cudaStreamSynchronize(h->stream);
  cudaFree(d_tmp);
  if (e != cudaSuccess) {
    h->err = std::string("cub scan: ") + cudaGetErrorString(e);
    return EMBA_E_CUDA;
  }
  return EMBA_OK;
}

static void free_state(StateSlot& s) {
  cudaFree(s.quat); cudaFree(s.Gx); cudaFree(s.Gy); cudaFree(s.G2); cudaFree(s.H3); cudaFree(s.Ktab);
  cudaFree(s.RotTab); cudaFree(s.JacTab); cudaFree(s.dp); cudaFree(s.e); cudaFree(s.pix); cudaFree(s.hist);
  s = StateSlot();
}

// ros::Duration(double) rounding of the half span, as the reference evaluates
// `t_batch_bgn + timespan * 0.5` (src/emba/model.cpp:116-119): Duration::operator*(double) is
// Duration(toSec()*scale); Duration(double d) is sec=floor(d), nsec=round((d-sec)*1e9).
// Plain IEEE double arithmetic on the host (no FMA contraction), so odd spans round like the CPU reference.
static int64_t batch_mid_time(int64_t t_bgn, int64_t t_end) {
  const int64_t span = t_end - t_bgn;
  int64_t sec = span / 1000000000LL;
  int64_t nsec = span % 1000000000LL;
  if (nsec < 0) { nsec += 1000000000LL; sec -= 1; }
  volatile double tosec = (double)sec + 1e-9 * (double)nsec;
  volatile double d = tosec * 0.5;
  const int64_t s = (int64_t)std::floor(d);
  volatile double frac = (d - (double)s) * 1e9;
  const int64_t ns = (int64_t)std::round(frac);
  return t_bgn + s * 1000000000LL + ns;
}

int rebuild_static(Handle* h);
void comm_destroy(Handle* h);
struct PoissonPlan;
void poisson_plan_destroy(PoissonPlan* p);

}  // namespace emba

using namespace emba;

extern "C" {

const char* emba_version(void) { return "emba_b200 0.1 sm_100a"; }

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
  if (cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_host, cudaEventDisableTiming);
  if (cudaMallocHost((void**)&h->h_pin, 1024 * sizeof(int64_t)) != cudaSuccess) { delete h; return EMBA_E_CUDA; }
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
  if (!ok) { emba_destroy((emba_handle_t)h); return EMBA_E_CUDA; }
  *out = (emba_handle_t)h;
  return EMBA_OK;
}

int emba_destroy(emba_handle_t hh) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_OK;
  cudaSetDevice(h->device);
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_destroy(h);
  if (h->poisson) { poisson_plan_destroy((PoissonPlan*)h->poisson); h->poisson = nullptr; }
  free_state(h->st[0]); free_state(h->st[1]);
  void* ptrs[] = {h->d_lut, h->d_tmid, h->d_spix_ev, h->d_pol, h->d_prev, h->d_refrank, h->d_bs, h->d_bu, h->d_rec, h->d_refpos,
                  h->d_items, h->d_gid, h->d_group_item0, h->d_part, h->d_scal, h->d_flags, h->d_amap, h->d_pflag, h->d_paidx, h->d_len, h->d_apix,
                  h->d_segoff, h->d_segend, h->d_gmask, h->d_gmask2, h->d_jrec, h->d_skey, h->d_sval, h->d_sval2, h->d_cub_tmp, h->d_sort_tmp, h->d_win64, h->d_winlo,
                  h->d_winhi, h->d_stripoff, h->d_strip, h->d_A22, h->d_b2, h->d_acc_part, h->d_gsum, h->d_A11,
                  h->d_b1, h->d_C, h->d_S, h->d_rhs, h->d_x1, h->d_x2, h->d_Spart, h->d_cg, h->d_ldlt_w, h->d_win2, h->d_win_all, h->d_own_len,
                  h->d_own_off, h->d_gwinlo, h->d_gwinhi, h->d_gstripoff, h->d_gstrip, h->d_recv};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {h->ev_fork, h->ev_join, h->ev_fork2, h->ev_join2, h->ev_sort0, h->ev_sort1, h->ev_host}) if (e) cudaEventDestroy(e);
  if (h->h_pin) cudaFreeHost(h->h_pin);
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
  if (rank != h->rank || world != h->world) { h->rank = rank; h->world = world; h->t0_ns = -1; }
  return EMBA_OK;
}

int emba_set_events(emba_handle_t hh, int64_t N, const uint16_t* x, const uint16_t* y, const int64_t* t_ns,
                    const uint8_t* pol) {
  Handle* h = (Handle*)hh;
  if (!h) return EMBA_E_ARG;
  if (N < 0 || (N > 0 && (!x || !y || !t_ns || !pol))) { h->err = "emba_set_events: null input"; return EMBA_E_ARG; }
  if (N >= (int64_t)1 << 31) { h->err = "emba_set_events: more than 2^31-1 events per window"; return EMBA_E_ARG; }
  EMBA_CUDA(cudaSetDevice(h->device));
  h->N = N;
  h->Nuse = (N / kBatch) * kBatch;  // integer division at model.cpp:79 drops the tail batch
  h->B = h->Nuse / kBatch;
  h->t0_ns = -1;  // forces the spline-dependent structures to be rebuilt
  h->st[0].evaluated = h->st[1].evaluated = false;
  h->formed = h->solved = false;
  h->Mc_total = 0;
  h->Mc = 0;
  const int64_t Nu = h->Nuse;
  // batch mid-times on the host (needs only two timestamps per batch)
  h->h_tmid.resize(h->B);
  for (int64_t b = 0; b < h->B; b++) h->h_tmid[b] = batch_mid_time(t_ns[b * kBatch], t_ns[b * kBatch + kBatch - 1]);
  EMBA_TRY(dev_alloc(h, &h->d_tmid, h->B));
  if (h->B) EMBA_CUDA(cudaMemcpyAsync(h->d_tmid, h->h_tmid.data(), sizeof(int64_t) * h->B, cudaMemcpyHostToDevice, h->stream));
  EMBA_TRY(dev_alloc(h, &h->d_spix_ev, Nu));
  EMBA_TRY(dev_alloc(h, &h->d_pol, Nu));
  EMBA_TRY(dev_alloc(h, &h->d_prev, Nu));
  EMBA_TRY(dev_alloc(h, &h->d_refrank, Nu));
  if (Nu == 0) { EMBA_CUDA(cudaStreamSynchronize(h->stream)); return EMBA_OK; }
  uint16_t *d_x = nullptr, *d_y = nullptr;
  uint32_t *d_ids = nullptr, *d_k2 = nullptr, *d_v2 = nullptr, *d_k1 = nullptr;
  int32_t *d_flag = nullptr, *d_rank = nullptr;
  int rc = EMBA_OK;
  do {
    if ((rc = dev_alloc(h, &d_x, Nu)) || (rc = dev_alloc(h, &d_y, Nu)) || (rc = dev_alloc(h, &d_ids, Nu)) ||
        (rc = dev_alloc(h, &d_k1, Nu)) || (rc = dev_alloc(h, &d_k2, Nu)) || (rc = dev_alloc(h, &d_v2, Nu)) ||
        (rc = dev_alloc(h, &d_flag, Nu)) || (rc = dev_alloc(h, &d_rank, Nu)))
      break;
    cudaMemcpyAsync(d_x, x, sizeof(uint16_t) * Nu, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(d_y, y, sizeof(uint16_t) * Nu, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(h->d_pol, pol, sizeof(uint8_t) * Nu, cudaMemcpyHostToDevice, h->stream);
    cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream);
    const int T = 256, G = ceil_div64(Nu, T);
    k_spix<<<G, T, 0, h->stream>>>(d_x, d_y, h->Ws, h->Hs, Nu, h->d_spix_ev, d_ids, h->d_flags);
    h->launches++;
    cudaMemcpyAsync(d_k1, h->d_spix_ev, sizeof(uint32_t) * Nu, cudaMemcpyDeviceToDevice, h->stream);
    uint32_t *ks = nullptr, *vs = nullptr;
    if ((rc = sort_pairs(h, d_k1, d_ids, d_k2, d_v2, Nu, bits_for((uint64_t)h->Ws * h->Hs), &ks, &vs))) break;
    k_pair_flags<<<G, T, 0, h->stream>>>(ks, Nu, d_flag);
    h->launches++;
    if ((rc = exclusive_sum(h, d_flag, d_rank, Nu))) break;
    k_pair_links<<<G, T, 0, h->stream>>>(vs, d_flag, d_rank, Nu, h->d_prev, h->d_refrank);
    h->launches++;
    int32_t last_rank = 0, last_flag = 0, flags0 = 0;
    cudaMemcpyAsync(&last_rank, d_rank + (Nu - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&last_flag, d_flag + (Nu - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&flags0, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { h->err = std::string("emba_set_events: ") + cudaGetErrorString(e); rc = EMBA_E_CUDA; break; }
    if (flags0 & 1) { h->err = "emba_set_events: event coordinates outside the sensor"; rc = EMBA_E_ARG; break; }
    h->Mc_total = (int64_t)last_rank + last_flag;
  } while (0);
  cudaFree(d_x); cudaFree(d_y); cudaFree(d_ids); cudaFree(d_k1); cudaFree(d_k2); cudaFree(d_v2); cudaFree(d_flag);
  cudaFree(d_rank);
  return rc;
}

int emba_num_pairs(emba_handle_t hh, int64_t* out) {
  Handle* h = (Handle*)hh;
  if (!h || !out) return EMBA_E_ARG;
  *out = h->Mc_total;
  return EMBA_OK;
}

}  // extern "C"

namespace emba {

// Rebuilds everything that depends on the spline time base (t0, dt, n) or on the shard: batch (s, u), the
// canonical measurement order with its records, groups and work items, and the per-measurement buffers.
int rebuild_static(Handle* h) {
  const int64_t Nu = h->Nuse, B = h->B;
  const int n = h->n;
  h->Mc = 0; h->n_items = 0; h->n_groups = 0; h->dmax = 0;
  h->h_items.clear();
  EMBA_TRY(dev_alloc(h, &h->d_bs, B));
  EMBA_TRY(dev_alloc(h, &h->d_bu, B));
  for (int s = 0; s < 2; s++) {
    EMBA_TRY(dev_alloc(h, &h->st[s].quat, (int64_t)n * 4));
    EMBA_TRY(dev_alloc(h, &h->st[s].Ktab, (int64_t)n * kKnotStride));
    EMBA_TRY(dev_alloc(h, &h->st[s].RotTab, B));
    EMBA_TRY(dev_alloc(h, &h->st[s].JacTab, B));
    h->st[s].evaluated = false;
  }
  h->formed = h->solved = false;
  EMBA_TRY(dev_alloc(h, &h->d_A11, (int64_t)9 * n * n));
  EMBA_TRY(dev_alloc(h, &h->d_b1, (int64_t)3 * n));
  EMBA_TRY(dev_alloc(h, &h->d_x1, (int64_t)3 * n));
  EMBA_TRY(dev_alloc(h, &h->d_S, (int64_t)(3 * n + 1) * (3 * n + 1)));  // bordered with the rhs row
  EMBA_TRY(dev_alloc(h, &h->d_rhs, (int64_t)3 * n));
  if (B == 0 || h->Mc_total == 0) {
    for (int s = 0; s < 2; s++) {
      EMBA_TRY(dev_alloc(h, &h->st[s].dp, 1)); EMBA_TRY(dev_alloc(h, &h->st[s].e, 1)); EMBA_TRY(dev_alloc(h, &h->st[s].pix, 1));
    }
    EMBA_TRY(dev_alloc(h, &h->d_rec, 1));
    EMBA_TRY(dev_alloc(h, &h->d_refpos, 1));
    EMBA_TRY(dev_alloc(h, &h->d_gid, (int64_t)n));
    EMBA_CUDA(cudaMemset(h->d_gid, 0xFF, sizeof(int32_t) * n));
    return EMBA_OK;
  }
  const int T = 256;
  EMBA_CUDA(cudaMemsetAsync(h->d_flags, 0, sizeof(int32_t) * 16, h->stream));
  k_batch_su<<<ceil_div64(B, T), T, 0, h->stream>>>(h->d_tmid, B, h->t0_ns, h->dt_ns, n, h->d_bs, h->d_bu, h->d_flags);
  EMBA_LAUNCH_CHECK();
  int32_t flags0 = 0;
  EMBA_CUDA(cudaMemcpyAsync(&flags0, h->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  EMBA_CUDA(cudaStreamSynchronize(h->stream));
  if (flags0 & 2) {
    h->err = "batch mid-time outside the spline support (the reference aborts here: so3_spline.h:221-230)";
    return EMBA_E_SUPPORT;
  }
  const int64_t Mt = h->Mc_total;
  int32_t *d_flag = nullptr, *d_pos = nullptr, *d_gidx = nullptr, *d_gstart = nullptr;
  uint32_t *d_key = nullptr, *d_ev = nullptr, *d_key2 = nullptr, *d_ev2 = nullptr, *d_gkey = nullptr;
  int rc = EMBA_OK;
  std::vector<int32_t> gstart;
  std::vector<uint32_t> gkey;
  uint32_t *ks = nullptr, *vs = nullptr;
  do {
    if ((rc = dev_alloc(h, &d_flag, std::max(Nu, Mt))) || (rc = dev_alloc(h, &d_pos, std::max(Nu, Mt))) ||
        (rc = dev_alloc(h, &d_key, Mt)) || (rc = dev_alloc(h, &d_ev, Mt)) || (rc = dev_alloc(h, &d_key2, Mt)) ||
        (rc = dev_alloc(h, &d_ev2, Mt)))
      break;
    k_meas_keys<<<ceil_div64(Nu, T), T, 0, h->stream>>>(h->d_prev, h->d_bs, Nu, n, d_flag);
    h->launches++;
    if ((rc = exclusive_sum(h, d_flag, d_pos, Nu))) break;
    k_meas_compact<<<ceil_div64(Nu, T), T, 0, h->stream>>>(h->d_prev, h->d_bs, d_pos, Nu, n, d_key, d_ev);
    h->launches++;
    if ((rc = sort_pairs(h, d_key, d_ev, d_key2, d_ev2, Mt, bits_for((uint64_t)n * n), &ks, &vs))) break;
    // group heads
    k_head_flags<<<ceil_div64(Mt, T), T, 0, h->stream>>>(ks, Mt, d_flag);
    h->launches++;
    if ((rc = exclusive_sum(h, d_flag, d_pos, Mt))) break;
    int32_t lastpos = 0, lastflag = 0;
    cudaMemcpyAsync(&lastpos, d_pos + (Mt - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(&lastflag, d_flag + (Mt - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "rebuild_static: sync failed"; rc = EMBA_E_CUDA; break; }
    const int G = lastpos + lastflag;
    if ((rc = dev_alloc(h, &d_gstart, G)) || (rc = dev_alloc(h, &d_gkey, G))) break;
    k_head_scatter<<<ceil_div64(Mt, T), T, 0, h->stream>>>(ks, d_flag, d_pos, Mt, d_gstart, d_gkey);
    h->launches++;
    gstart.resize(G); gkey.resize(G);
    cudaMemcpyAsync(gstart.data(), d_gstart, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(gkey.data(), d_gkey, sizeof(uint32_t) * G, cudaMemcpyDeviceToHost, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "rebuild_static: sync failed"; rc = EMBA_E_CUDA; break; }
  } while (0);
  if (rc == EMBA_OK) {
    // host: split groups into work items, pick this rank's contiguous slice (time sharding: groups are ordered by
    // cp_c, i.e. by time), build the lookup tables
    const int G = (int)gstart.size();
    std::vector<WorkItem> all;
    for (int g = 0; g < G; g++) {
      const int64_t s0 = gstart[g], s1 = (g + 1 < G) ? gstart[g + 1] : Mt;
      const int cc = (int)(gkey[g] / (uint32_t)n), cp = (int)(gkey[g] % (uint32_t)n);
      h->dmax = std::max(h->dmax, cc - cp);
      static const int64_t item_max = getenv("EMBA_ITEM_MAX") ? atoll(getenv("EMBA_ITEM_MAX")) : kItemMax;
      for (int64_t s = s0; s < s1; s += item_max) {
        WorkItem w;
        w.cp_c = cc; w.cp_p = cp; w.start = (int32_t)s; w.count = (int32_t)std::min<int64_t>(item_max, s1 - s); w.group = g;
        all.push_back(w);
      }
    }
    // shard boundaries in measurements, snapped to item starts
    const int64_t lo_t = Mt * h->rank / h->world, hi_t = Mt * (h->rank + 1) / h->world;
    int64_t m_lo = -1, m_hi = -1;
    for (const WorkItem& w : all) {
      if (w.start >= lo_t && w.start < hi_t) {
        if (m_lo < 0) m_lo = w.start;
        m_hi = (int64_t)w.start + w.count;
        h->h_items.push_back(w);
      }
    }
    if (m_lo < 0) { m_lo = m_hi = 0; }
    h->Mc = m_hi - m_lo;
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
    h->iota_len = 0;
    std::vector<int32_t> gid((size_t)n * (h->dmax + 1), -1);
    for (const WorkItem& w : h->h_items) gid[(size_t)w.cp_c * (h->dmax + 1) + (w.cp_c - w.cp_p)] = w.group;
    do {
      if ((rc = dev_alloc(h, &h->d_items, h->n_items)) || (rc = dev_alloc(h, &h->d_gid, (int64_t)gid.size())) ||
          (rc = dev_alloc(h, &h->d_group_item0, (int64_t)item0.size())) || (rc = dev_alloc(h, &h->d_rec, h->Mc)) ||
          (rc = dev_alloc(h, &h->d_refpos, h->Mc)) ||
          (rc = dev_alloc(h, &h->d_acc_part, (int64_t)h->n_items * kAccN)) ||
          (rc = dev_alloc(h, &h->d_skey, h->Mc)) || (rc = dev_alloc(h, &h->d_sval, h->Mc)) ||
          (rc = dev_alloc(h, &h->d_sval2, h->Mc)) ||
          (rc = dev_reserve(h, &h->d_jrec, &h->jrec_cap, h->Mc * kRecDoubles)) ||
          (rc = dev_alloc(h, &h->d_gsum, (int64_t)h->n_groups * kAccN)))
        break;
      for (int s = 0; s < 2 && rc == EMBA_OK; s++) {
        if ((rc = dev_alloc(h, &h->st[s].dp, h->Mc)) || (rc = dev_alloc(h, &h->st[s].e, h->Mc)) ||
            (rc = dev_alloc(h, &h->st[s].pix, h->Mc)))
          break;
      }
      if (rc) break;
      if (h->n_items) cudaMemcpyAsync(h->d_items, h->h_items.data(), sizeof(WorkItem) * h->n_items, cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(h->d_gid, gid.data(), sizeof(int32_t) * gid.size(), cudaMemcpyHostToDevice, h->stream);
      cudaMemcpyAsync(h->d_group_item0, item0.data(), sizeof(int32_t) * item0.size(), cudaMemcpyHostToDevice, h->stream);
      if (h->Mc) {
        k_build_recs<<<ceil_div64(h->Mc, T), T, 0, h->stream>>>(vs, m_lo, h->Mc, h->d_spix_ev, h->d_pol, h->d_prev,
                                                                h->d_refrank, h->d_lut, h->d_rec, h->d_refpos);
        h->launches++;
      }
      cudaError_t e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { h->err = std::string("rebuild_static: ") + cudaGetErrorString(e); rc = EMBA_E_CUDA; }
    } while (0);
  }
  cudaFree(d_flag); cudaFree(d_pos); cudaFree(d_key); cudaFree(d_ev); cudaFree(d_key2); cudaFree(d_ev2);
  cudaFree(d_gstart); cudaFree(d_gkey); cudaFree(d_gidx);
  return rc;
}

}  // namespace emba
