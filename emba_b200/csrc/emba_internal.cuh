// emba_b200 internal declarations: handle layout, error macros, device math.
// All arithmetic on the measurement path is fp64 (the reference is fp64 throughout and its
// cost function is discontinuous: SURVEY.md "five facts" #4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/emba_b200.h"

namespace emba {

constexpr int kBatch = 100;         // hard-coded event batch size (reference src/emba/model.cpp:78)
constexpr int kKnotStride = 12;     // doubles per knot-interval entry: R_s (9) + delta_w (3)
constexpr int kItemMax = 4096;      // measurements per pose-block assembly work item (16 tiles): measured best for k_asm_pose on C2 (2048: 0.475, 4096: 0.481, 8192: 0.502, 16384: 0.616 ms); EMBA_ITEM_MAX overrides
constexpr int kAccN = 91;           // upper triangle of the 13x13 outer product of [Jc Jp e]
constexpr int kRecDoubles = 16;     // Jacobian-row record: Jc[6] Jp[6] e dp[2] meta  (128 bytes)
constexpr double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

// static per-measurement record, canonical order (sorted by (cp_c, cp_p), then time of the current event)
struct __align__(32) MeasRec {
  double bx, by, bz;  // bearing vector of the sensor pixel (both events of a pair share it): saves a LUT gather
  uint32_t bc_pol;    // batch of the current event | polarity << 31
  uint32_t bp;        // batch of the previous event
};

struct WorkItem {
  int32_t cp_c, cp_p;  // control-pose indices of the group
  int32_t start, count;  // measurement range in canonical order (local to the shard)
  int32_t group;       // group id
};

// One grow-only device allocation carved into 256-byte aligned pieces: the per-window structures are laid out in
// arenas, so a window of the size of the previous one allocates nothing (and the first one calls cudaMalloc once
// per arena, not once per array).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void reset() { off = 0; }
  static size_t pad(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  template <typename T>
  T* take(int64_t count) {
    const size_t bytes = pad(sizeof(T) * (size_t)(count > 0 ? count : 1));
    if (off + bytes > cap) return nullptr;
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

// host -> device copies of caller memory: direct when the source is page-locked, otherwise staged through two pinned
// 8 MB buffers (memcpy of chunk k+1 overlaps the DMA of chunk k)
struct Uploader {
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int turn = 0;
};
constexpr size_t kStageBytes = (size_t)8 << 20;

// device-resident event sequence (include/emba_b200.h: emba_events_t)
struct EventStore {
  int device = 0;
  int64_t N = 0, cap = 0;
  uint16_t *x = nullptr, *y = nullptr;
  int64_t* t = nullptr;
  uint8_t* pol = nullptr;
  cudaStream_t stream = nullptr;
  Uploader up;
  int64_t* h_pin = nullptr;
  std::string err;
};

// device-resident optimisation state + outputs of its last evaluation
struct StateSlot {
  double* quat = nullptr;    // [n*4] xyzw
  double* Gx = nullptr;      // [P]
  double* Gy = nullptr;      // [P]
  double2* G2 = nullptr;     // [P] interleaved (Gx, Gy)
  double4* H3 = nullptr;     // [P] (Gxx, Gxy, Gyy, 0)
  double* Ktab = nullptr;    // [n*kKnotStride] per knot interval s: R_s (9, row-major), delta_w = Log(R_{s+1} R_s^-1) (3)
  double4* RotTab = nullptr; // [B] per batch: (s1, s2, knot index s, 0): R_batch = (I + s1 K + s2 K^2) R_s, K = [delta_w]x
  double4* JacTab = nullptr; // [B] per batch: (alpha, beta, gamma, 0): A = alpha I + beta K + gamma K^2
  double2* dp = nullptr;     // [Mc] displacement pm_c - pm_p
  double* e = nullptr;       // [Mc] residual
  int32_t* pix = nullptr;    // [Mc] pano pixel index of the current event, -1 = outlier
  int32_t* slot = nullptr;   // [Mc] arrival rank of the measurement among the (local) rows of its pixel: the value the
                             //      histogram atomic returns; unique per pixel, arbitrary order
  int32_t* hist = nullptr;   // [P] num_ev_map (global: all-reduced over the ranks)
  int32_t* hist_loc = nullptr;  // [P] this rank's own counts (== hist with one GPU)
  double cost_data = 0, cost_reg = 0;
  int64_t M = 0;
  bool evaluated = false;
};

struct Handle {
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: the row sort runs beside the pose-side assembly
  cudaStream_t stream3 = nullptr;  // communication stream of the pipelined strip exchange (several GPUs)
  cudaEvent_t ev_chunk[64] = {};   // k_pix finished the pixel range of owner (rank + s)
  cudaEvent_t ev_comm = nullptr;
  cudaEvent_t ev_x[8] = {};        // timing marks of the multi-GPU exchange phases
  double t_comm_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<int64_t> x_recv_cnt, x_send_off, x_recvbase;  // strip exchange: poses received per source rank, my
  int64_t x_gtot = 0;                                        // strip offsets at the ownership boundaries, merged total
  int64_t* d_glen = nullptr;       // [P+2] merged strip lengths
  // peer-memory strip exchange (comm.cu): the map-side kernel stores every finished sub-strip straight into its
  // owner's receive buffer over NVLink (CUDA IPC mapping of every rank's d_recv)
  static constexpr int kPeerMax = 16;
  int peer_mode = -1;                          // -1 not negotiated yet, 0 off (ncclSend/ncclRecv all-to-all), 1 on
  bool peer_now = false;                       // this assembly's strips go through peer memory
  double* peer_recv[kPeerMax] = {};            // every rank's receive buffer as mapped here (own entry = d_recv)
  void* peer_base[kPeerMax] = {};              // what cudaIpcOpenMemHandle returned (closed on re-map / destroy)
  int64_t peer_cap[kPeerMax] = {};             // capacity (doubles) of every rank's buffer, tracked identically everywhere
  std::vector<int64_t> x_cnt;                  // [source][owner] poses of sub-strips (identical on all ranks)
  int64_t* d_dst = nullptr;                    // [Np] destination address of every local sub-strip
  int64_t dst_cap = 0;
  int64_t* d_peerx = nullptr;                  // device staging of the handle all-gather
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_sort0 = nullptr, ev_sort1 = nullptr;
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev_host = nullptr;   // marks a small device->host read-back the host waits for while later launches queue
  int64_t* h_pin = nullptr;        // pinned host scratch (1024 x int64) for those read-backs
  int sm_count = 148;
  int64_t launches = 0;
  Arena ar_ev, ar_meas, ar_tmp;    // per-window event-level / measurement-level structures, scratch
  Uploader up;
  int strict_range = 0;            // emba_set_strict_range
  double t_setup_ms[2] = {0, 0};   // device time of the last emba_set_events / static rebuild
  // config
  int Ws = 0, Hs = 0, Wp = 0, Hp = 0;
  int64_t P = 0;
  double C_th = 0;
  double* d_lut = nullptr;  // [Ws*Hs*3]
  // events
  int64_t N = 0, Nuse = 0, B = 0;  // events of the window, the usable multiple of 100, batches (all GLOBAL)
  int64_t ev_off = 0, Nloc = 0;    // this rank's slice of the events (whole batches): first global id, count. The
                                   // per-event arrays below are local (index = global id - ev_off); d_prev holds
                                   // GLOBAL ids. One GPU: ev_off = 0, Nloc = Nuse.
  int64_t Mloc_pairs = 0;          // pairs whose current event lies in the slice
  int64_t* d_tmid = nullptr;     // [B]
  uint32_t* d_spix_ev = nullptr; // [Nuse] sensor pixel per event
  uint8_t* d_pol = nullptr;      // [Nuse]
  int32_t* d_prev = nullptr;     // [Nuse] previous event at the same sensor pixel, -1 if none
  uint32_t* d_refrank = nullptr; // [Nuse] rank of the pair (as current event) in reference order
  int64_t Mc_total = 0;          // pairs in the window
  // spline time base and everything that depends on it
  int64_t t0_ns = -1, dt_ns = -1;
  int n = 0;
  int32_t* d_bs = nullptr;    // [B] knot index s of the batch mid-time
  double* d_bu = nullptr;     // [B] u in [0,1)
  MeasRec* d_rec = nullptr;   // [Mc] canonical order, this shard only (32 B, read as two coalesced 16 B loads)
  uint32_t* d_refpos = nullptr;  // [Mc] rank of the pair in the reference's order (sensor pixel row-major, then time)
  int64_t Mc = 0;             // measurements of this shard
  std::vector<WorkItem> h_items;
  WorkItem* d_items = nullptr;
  int n_items = 0, n_groups = 0, dmax = 0;
  int32_t* d_gid = nullptr;         // [n*(dmax+1)] group id of (cp_c, d) or -1
  int32_t* d_group_item0 = nullptr; // [n_groups+1] first item of each group
  // shard
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  // states
  StateSlot st[2];
  int cur = 0;  // index of the CURRENT slot; candidate = 1-cur
  // scratch for reductions
  double* d_part = nullptr;  // per-block partial sums
  int64_t part_cap = 0;
  double* d_scal = nullptr;  // small device scalars
  int32_t* d_flags = nullptr;  // error flags
  // normal equations (device)
  int thres = 0;
  int64_t Np = 0;
  int32_t* d_amap = nullptr;    // [P] active index or -1
  int32_t* d_pflag = nullptr;   // [P] scratch: active flag
  int32_t* d_paidx = nullptr;   // [P] scratch: exclusive scan of the flags
  int64_t* d_len = nullptr;     // [P+1] scratch: strip lengths
  int32_t* d_apix = nullptr;    // [Np] pixel index of each active pixel
  int32_t* d_segoff = nullptr;  // [Np] first row of each active pixel's segment in the sorted row list
  int32_t* d_segend = nullptr;  // [Np] one past its last row
  int64_t Ma = 0;               // measurements on active pixels
  double* d_jrec = nullptr;     // [Mc*16] Jacobian rows
  int64_t jrec_cap = 0;
  // TMA descriptor of d_jrec as a [Mc][16] fp64 tensor, 256-row boxes, 128-byte swizzle (k_asm_pose stores)
  alignas(64) unsigned char jrec_tmap[128] = {};
  const void* jrec_tmap_ptr = nullptr;
  int64_t jrec_tmap_rows = 0;
  uint32_t* d_sval = nullptr;   // [Mc] row ids grouped by active pixel, ascending inside a pixel
  int32_t* d_segcnt = nullptr;  // [P+1] scratch: local rows of each pixel if it is active, else 0
  int32_t* d_longlist = nullptr;  // [P+1] active pixels whose segment is too long for the warp sort; [0] = count
  void* d_scan_tmp = nullptr;   // block sums of the scans over the panorama
  bool jrec_valid = false;      // d_jrec holds the rows of the last assembly (it doubles as scratch of downloads)
  int2* d_win64 = nullptr;      // [Np] (first, last) control pose touching the pixel while k_asm_pose accumulates them
  int32_t* d_winlo = nullptr;   // [Np] first control pose touching the pixel
  int32_t* d_winhi = nullptr;   // [Np] last control pose touching the pixel
  int64_t* d_stripoff = nullptr;  // [Np+1] offsets (in poses) of the per-pixel A12 strips
  int64_t strip_total = 0;
  double* d_strip = nullptr;    // [strip_total*6] A12, per pixel per pose: 3 rows x 2 cols
  int64_t strip_cap = 0;
  // solve view of A12: with one GPU it aliases the arrays above; with several GPUs every rank owns a contiguous
  // range of active pixels and holds their complete strips (merged from all time slices), the other pixels have
  // empty windows (lo > hi)
  // coarse occupancy of each pixel's strip: bit k set when some entry of pose group k (pose_group poses per
  // group, at most 64 groups) is non-zero. Revisiting trajectories leave long windows mostly empty; the Schur
  // kernel skips (pixel, tile pair) products whose tiles hold no entries.
  unsigned long long* d_gmask = nullptr;   // [2(P+1)] local / single-GPU: (fix = 0, fix = 1) per pixel
  unsigned long long* d_gmask2 = nullptr;  // [2(P+1)] merged strips (multi-GPU owner view)
  unsigned long long* sv_gmask = nullptr;
  int pose_group = 16;
  int mask_min_len = -1;  // >= 0: masks of strips at least this long are still to be computed (by the solve)
  int32_t* sv_winlo = nullptr;
  int32_t* sv_winhi = nullptr;
  int64_t* sv_stripoff = nullptr;
  double* sv_strip = nullptr;
  int64_t sv_strip_total = 0;
  int32_t* d_win2 = nullptr;       // [Np*2] packed local windows
  int32_t* d_win_all = nullptr;    // [world][Np*2] windows of every rank
  int64_t win_all_cap = 0;
  int64_t* d_own_len = nullptr;    // [world][n_own+1] sub-strip lengths of my pixels per source rank
  int64_t* d_own_off = nullptr;    // [world][n_own+1] exclusive scans of the above
  int64_t own_cap = 0;
  int32_t* d_gwinlo = nullptr;     // [Np] merged windows (owned pixels), empty elsewhere
  int32_t* d_gwinhi = nullptr;
  int64_t* d_gstripoff = nullptr;  // [Np+1]
  double* d_gstrip = nullptr;
  int64_t gstrip_cap = 0;
  double* d_recv = nullptr;        // received sub-strips, one contiguous chunk per source rank
  int64_t recv_cap = 0;
  double* d_A22 = nullptr;      // [Np*3] xx, xy, yy
  double* d_b2 = nullptr;       // [Np*2]
  double* d_acc_part = nullptr; // [n_items*91]
  double* d_gsum = nullptr;     // [n_groups*91]
  double* d_A11 = nullptr;      // [3n*3n]
  double* d_b1 = nullptr;       // [3n]
  int64_t A11_cap = 0, b1_cap = 0, x1_cap = 0, S_cap = 0, rhs_cap = 0;
  bool formed = false;
  int map_path = 0;  // EMBA_MAP_SORTED / EMBA_MAP_ATOMIC
  bool a11_partial = false;  // multi-GPU: A11/b1 hold this rank's partial sums (combined inside the Schur all-reduce)
  // solve
  double* d_C = nullptr;        // [Np*3] inverse of damped A22
  double* d_S = nullptr;        // [d*d]
  double* d_rhs = nullptr;      // [d]
  double* d_x1 = nullptr;       // [3n]
  double* d_x2 = nullptr;       // [2Np]
  double* d_ldlt_w = nullptr;   // [d*32] panel scratch of the blocked LDL^T
  int64_t ldlt_w_cap = 0;
  double* d_Spart = nullptr;
  int64_t Spart_cap = 0;
  double* d_cg = nullptr;       // PCG vectors
  int64_t cg_cap = 0;
  int solved_fix = 0;
  bool solved = false;
  void* poisson = nullptr;  // lazily created PoissonPlan of emba_reconstruct_map (poisson.cu)
  // timing
  cudaEvent_t ev[12];
  double t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

#define EMBA_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
               std::to_string(__LINE__);                                                    \
      return EMBA_E_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define EMBA_TRY(call)        \
  do {                        \
    int _r = (call);          \
    if (_r != EMBA_OK) return _r; \
  } while (0)

#define EMBA_LAUNCH_CHECK()      \
  do {                           \
    h->launches++;               \
    EMBA_CUDA(cudaGetLastError()); \
  } while (0)

// (re)allocation: the new buffer is obtained first, so a failure leaves the old one (and the handle) intact
template <typename T>
inline int dev_alloc(Handle* h, T** p, int64_t count) {
  if (count <= 0) count = 1;
  T* q = nullptr;
  cudaError_t e = cudaMalloc((void**)&q, sizeof(T) * (size_t)count);
  if (e != cudaSuccess && *p) {  // not enough room for both: give the old one back and retry once
    cudaGetLastError();
    cudaFree(*p);
    *p = nullptr;
    e = cudaMalloc((void**)&q, sizeof(T) * (size_t)count);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    h->err = std::string("cudaMalloc of ") + std::to_string(sizeof(T) * (size_t)count) + " bytes: " + cudaGetErrorString(e);
    return EMBA_E_CUDA;
  }
  if (*p) cudaFree(*p);
  *p = q;
  return EMBA_OK;
}
inline int arena_reserve(Handle* h, Arena& a, size_t bytes) {
  a.off = 0;
  if (a.base && a.cap >= bytes) return EMBA_OK;
  const size_t want = bytes + bytes / 16 + 4096;
  char* q = nullptr;
  if (a.base) { cudaFree(a.base); a.base = nullptr; a.cap = 0; }
  cudaError_t e = cudaMalloc((void**)&q, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    h->err = std::string("cudaMalloc of ") + std::to_string(want) + " bytes (arena): " + cudaGetErrorString(e);
    return EMBA_E_CUDA;
  }
  a.base = q;
  a.cap = want;
  return EMBA_OK;
}
template <typename T>
inline int dev_reserve(Handle* h, T** p, int64_t* cap, int64_t count) {
  if (*p && *cap >= count) return EMBA_OK;
  int64_t want = count + count / 8 + 16;
  EMBA_TRY(dev_alloc(h, p, want));
  *cap = want;
  return EMBA_OK;
}

inline int ceil_div64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// prims.cu: device-wide exclusive scan and stable radix sort (hand-written; the library uses no CUB)
size_t scan_scratch_bytes(int64_t count);
template <typename T>
int scan_exclusive(Handle* h, cudaStream_t st, const T* in, T* out, int64_t count, void* scratch);
bool is_pinned(const void* p);
cudaError_t uploader_init(Uploader& up);
cudaError_t upload_bytes(Uploader& up, cudaStream_t st, void* dst, const void* src, size_t bytes);
cudaError_t download_bytes(Uploader& up, cudaStream_t st, void* dst, const void* src, size_t bytes);
void uploader_free(Uploader& up);
size_t radix_scratch_bytes(int64_t count);
int radix_sort_pairs(Handle* h, cudaStream_t st, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, int64_t count,
                     int nbits, void* scratch, bool iota, bool keep_keys, int* which);

// ------------------------------------------------------------------------------------------------
// device math
// ------------------------------------------------------------------------------------------------
struct Vec3 { double x, y, z; };
struct Mat3 { double m[9]; };  // row-major

// Hamilton product, xyzw (Sophus SO3::operator*)
__device__ __forceinline__ double4 quat_mul(const double4& a, const double4& b) {
  double4 r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}

// Eigen::Quaternion::toRotationMatrix
__device__ __forceinline__ Mat3 quat_to_R(const double4& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 R = {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
             1 - (txx + tyy)}};
  return R;
}

// Sophus SO3::logAndTheta (reference thirdparty/basalt-headers/thirdparty/Sophus/sophus/so3.hpp:247-290)
__device__ __forceinline__ Vec3 so3_log(const double4& q) {
  const double sq = q.x * q.x + q.y * q.y + q.z * q.z;
  const double w = q.w;
  double f;
  if (sq < kSophusEps * kSophusEps) {
    f = 2.0 / w - (2.0 / 3.0) * sq / (w * w * w);
  } else {
    const double n = sqrt(sq);
    if (fabs(w) < kSophusEps) f = (w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = 2.0 * atan(n / w) / n;
  }
  Vec3 r = {f * q.x, f * q.y, f * q.z};
  return r;
}

// Sophus SO3::expAndTheta (so3.hpp:583-619)
__device__ __forceinline__ double4 so3_exp(const Vec3& o) {
  const double th2 = o.x * o.x + o.y * o.y + o.z * o.z;
  double imag, real;
  if (th2 < kSophusEps * kSophusEps) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    const double th = sqrt(th2);
    const double half = 0.5 * th;
    imag = sin(half) / th;
    real = cos(half);
  }
  double4 q = {imag * o.x, imag * o.y, imag * o.z, real};
  return q;
}

// 256-bit read-only global load (sm_100: LDG.E.ENL2.256): one instruction per 32-byte table entry / record, half
// the L1TEX wavefronts of two 128-bit gathers. The address must be 32-byte aligned.
__device__ __forceinline__ double4 ldg256(const void* p) {
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}

// Batch pose applied to a bearing vector. The linear SO(3) spline gives R_batch = Exp(u delta_w) R_s with
// delta_w = Log(R_{s+1} R_s^-1) (so3_spline.h:218-274 in world-frame form), i.e. Rodrigues with the per-batch
// scalars s1 = sin(u th)/th, s2 = (1 - cos(u th))/th^2 and the per-knot-interval matrix K = [delta_w]x:
//   rb = r0 + s1 (delta x r0) + s2 (delta x (delta x r0)),  r0 = R_s b
__device__ __forceinline__ void rotate_bearing(const double* __restrict__ kt, double s1, double s2, double bx,
                                               double by, double bz, double& X, double& Y, double& Z) {
  const double x0 = kt[0] * bx + kt[1] * by + kt[2] * bz;
  const double y0 = kt[3] * bx + kt[4] * by + kt[5] * bz;
  const double z0 = kt[6] * bx + kt[7] * by + kt[8] * bz;
  const double dx = kt[9], dy = kt[10], dz = kt[11];
  const double x1 = dy * z0 - dz * y0, y1 = dz * x0 - dx * z0, z1 = dx * y0 - dy * x0;
  const double x2 = dy * z1 - dz * y1, y2 = dz * x1 - dx * z1, z2 = dx * y1 - dy * x1;
  X = x0 + s1 * x1 + s2 * x2;
  Y = y0 + s1 * y1 + s2 * y2;
  Z = z0 + s1 * z1 + s2 * z2;
}

// the same with the knot-interval entry fetched as three 256-bit loads (the entry is 96 bytes, 32-byte aligned): 3 LSU
// instructions instead of 12 scalar ones -- k_eval was bound by L1TEX wavefronts
__device__ __forceinline__ void rotate_bearing_v(const double* __restrict__ kt, double s1, double s2, double bx,
                                                 double by, double bz, double& X, double& Y, double& Z) {
  const double4 a = ldg256(kt), b = ldg256(kt + 4), c = ldg256(kt + 8);
  const double x0 = a.x * bx + a.y * by + a.z * bz;
  const double y0 = a.w * bx + b.x * by + b.y * bz;
  const double z0 = b.z * bx + b.w * by + c.x * bz;
  const double dx = c.y, dy = c.z, dz = c.w;
  const double x1 = dy * z0 - dz * y0, y1 = dz * x0 - dx * z0, z1 = dx * y0 - dy * x0;
  const double x2 = dy * z1 - dz * y1, y2 = dz * x1 - dx * z1, z2 = dx * y1 - dy * x1;
  X = x0 + s1 * x1 + s2 * x2;
  Y = y0 + s1 * y1 + s2 * y2;
  Z = z0 + s1 * z1 + s2 * z2;
}

// w = v A for a row vector v, A = alpha I + beta K + gamma K^2, K = [delta]x:  v K = v x delta
__device__ __forceinline__ void row_times_A(const double* __restrict__ kt, double al, double be, double ga,
                                            const double v[3], double w[3]) {
  const double dx = kt[9], dy = kt[10], dz = kt[11];
  const double a0 = v[1] * dz - v[2] * dy, a1 = v[2] * dx - v[0] * dz, a2 = v[0] * dy - v[1] * dx;
  const double b0 = a1 * dz - a2 * dy, b1 = a2 * dx - a0 * dz, b2 = a0 * dy - a1 * dx;
  w[0] = al * v[0] + be * a0 + ga * b0;
  w[1] = al * v[1] + be * a1 + ga * b1;
  w[2] = al * v[2] + be * a2 + ga * b2;
}

struct PanoCam {
  double fx, fy, cx, cy;  // fx = W/(2 pi), fy = H/pi, centre (W/2, H/2) (equirectangular_camera.h:11-16,64-67)
};

// EquirectangularCamera::projectToImage without the Jacobian (include/utils/equirectangular_camera.h:18-45)
__device__ __forceinline__ void project_pm(const PanoCam& c, double X, double Y, double Z, double& px, double& py) {
  const double phi = atan2(X, Z);
  const double rho = sqrt(X * X + Y * Y + Z * Z);
  const double theta = asin(Y / rho);
  px = c.cx + phi * c.fx;
  py = c.cy + theta * c.fy;
}

// the same for a UNIT bearing (records hold normalised bearings, rotations keep the norm to rounding): no norm, no
// division; the clamp only guards asin against |y| = 1 + 1 ulp at the poles
__device__ __forceinline__ void project_pm_unit(const PanoCam& c, double X, double Y, double Z, double& px, double& py) {
  const double phi = atan2(X, Z);
  const double theta = asin(fmin(1.0, fmax(-1.0, Y)));
  px = c.cx + phi * c.fx;
  py = c.cy + theta * c.fy;
}

// M = dpm_drb * drb_ddrot (2x3): projection Jacobian (equirectangular_camera.h:31-43) times -[rb]x
// (src/utils/event_pano_warper.cpp:62-65). With rb = (X, Y, Z), s = X^2 + Z^2 the product of the reference's
// two matrices simplifies exactly to
//   [ -fx X Y / s,  fx,  -fx Y Z / s ]
//   [ -fy Z / sqrt(s),  0,  fy X / sqrt(s) ]
// (the norm of rb cancels), which needs one reciprocal square root and no division. M[4] == 0 is dropped.
__device__ __forceinline__ void project_jac(const PanoCam& c, double X, double Y, double Z, double M[6]) {
  const double i1 = rsqrt(X * X + Z * Z);
  const double i2 = i1 * i1;
  const double fxy = c.fx * Y * i2;
  M[0] = -fxy * X;
  M[1] = c.fx;
  M[2] = -fxy * Z;
  M[3] = -c.fy * Z * i1;
  M[4] = 0.0;
  M[5] = c.fy * X * i1;
}

// occupancy masks of one strip (len poses from pose lo, 6 doubles per pose) computed by a warp; every lane returns
// them. Two masks, one per gauge choice of the solve (fix = 0 / 1 drops the first pose, which shifts the 16-pose
// Schur tiles by one pose): bit k of m[fix] <-> poses [group*k + fix, group*(k+1) + fix)
__device__ __forceinline__ void strip_mask_warp(const double* sp, int len, int lo, int group, int lane,
                                                unsigned long long& m0, unsigned long long& m1) {
  m0 = 0ull; m1 = 0ull;
  for (int i = lane; i < len * 6; i += 32)
    if (sp[i] != 0.0) {
      const int pose = lo + i / 6;
      m0 |= 1ull << min(63, pose / group);
      if (pose >= 1) m1 |= 1ull << min(63, (pose - 1) / group);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 |= __shfl_xor_sync(0xffffffffu, m0, o);
    m1 |= __shfl_xor_sync(0xffffffffu, m1, o);
  }
}

// index into the packed upper triangle of a symmetric 13x13 (i <= j)
__host__ __device__ __forceinline__ int tri13(int i, int j) { return i * 13 - (i * (i - 1)) / 2 + (j - i); }

}  // namespace emba
