/* A plain-C consumer of the emba_b200 C ABI (no Python, no torch, no C++): reads a scene from flat binary files,
 * runs evaluate -> form -> solve -> the whole LM loop, prints the numbers tests/test_cabi_consumer.py compares with
 * the golden outputs of the reference.
 *   usage: cabi_consumer <events.bin> <state.bin>
 * events.bin: EMBAEV01 (emba_b200/eventio.py). state.bin: int32 sensor_w, sensor_h, pano_w, pano_h, n_poses, pad;
 * double C_th, fx, fy, cx, cy, t_beg, dt_knots; double quat[4n]; double Gx[P]; double Gy[P]. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "emba_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != EMBA_OK) {                                                        \
      fprintf(stderr, "%s failed: %d %s\n", #call, rc_, emba_last_error(h));     \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

static void* slurp(const char* path, size_t* size) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc((size_t)n);
  if (fread(p, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(p); return NULL; }
  fclose(f);
  *size = (size_t)n;
  return p;
}

int main(int argc, char** argv) {
  emba_handle_t h = NULL;
  if (argc < 3) return 1;
  size_t esz = 0, ssz = 0;
  unsigned char* ev = (unsigned char*)slurp(argv[1], &esz);
  unsigned char* st = (unsigned char*)slurp(argv[2], &ssz);
  if (!ev || !st || memcmp(ev, "EMBAEV01", 8) != 0) { fprintf(stderr, "bad input files\n"); return 1; }
  int64_t N;
  memcpy(&N, ev + 8, 8);
  const uint16_t* x = (const uint16_t*)(ev + 32);
  const uint16_t* y = x + N;
  const uint8_t* pol = (const uint8_t*)(y + N);
  const int64_t* t_ns = (const int64_t*)(ev + 32 + 5 * N + ((8 - (5 * N) % 8) % 8));
  int32_t hdr[6];
  memcpy(hdr, st, sizeof(hdr));
  const int sw = hdr[0], sh = hdr[1], pw = hdr[2], ph = hdr[3], n = hdr[4];
  double par[7];
  memcpy(par, st + 24, sizeof(par));
  const double C_th = par[0], fx = par[1], fy = par[2], cx = par[3], cy = par[4], t_beg = par[5], dt = par[6];
  const double* quat = (const double*)(st + 24 + 56);
  const size_t P = (size_t)pw * ph;
  const double* Gx = quat + 4 * n;
  const double* Gy = Gx + P;
  /* bearing LUT of a zero-distortion pinhole camera (event_pano_warper.cpp:27-41) */
  double* lut = (double*)malloc(sizeof(double) * 3 * sw * sh);
  for (int v = 0; v < sh; v++)
    for (int u = 0; u < sw; u++) {
      double* b = lut + 3 * ((size_t)v * sw + u);
      b[0] = (u - cx) / fx; b[1] = (v - cy) / fy; b[2] = 1.0;
    }
  emba_config_t cfg;
  cfg.sensor_w = sw; cfg.sensor_h = sh; cfg.pano_w = pw; cfg.pano_h = ph; cfg.C_th = C_th; cfg.bearing_lut = lut;
  cfg.device = 0;
  int rc = emba_create(&cfg, &h);
  if (rc != EMBA_OK) { fprintf(stderr, "emba_create failed: %d (no CUDA device? there is no CPU fallback)\n", rc); return 3; }
  CHECK(emba_set_events(h, N, x, y, t_ns, pol));
  const int64_t t0_ns = (int64_t)(1e9 * t_beg), dt_ns = (int64_t)(1e9 * dt); /* trajectory.cpp:61-70 */
  CHECK(emba_set_state(h, EMBA_STATE_CURRENT, t0_ns, dt_ns, n, quat, Gx, Gy));
  double cd = 0, cr = 0;
  int64_t M = 0, Np = 0;
  CHECK(emba_evaluate(h, EMBA_STATE_CURRENT, EMBA_COST_QUADRATIC, 1.0, 5.0, &cd, &cr, &M));
  CHECK(emba_form_normal_eq(h, 5, EMBA_COST_QUADRATIC, 1.0, 5.0, &Np));
  double* x1 = (double*)malloc(sizeof(double) * 3 * n);
  double* x2 = (double*)malloc(sizeof(double) * 2 * (size_t)Np);
  CHECK(emba_solve(h, 1e-3, 0, 1, x1, x2, NULL, NULL));
  double s1 = 0, s2 = 0;
  for (int i = 0; i < 3 * (n - 1); i++) s1 += x1[i] * x1[i];
  for (int64_t i = 0; i < 2 * Np; i++) s2 += x2[i] * x2[i];
  printf("version %s\n", emba_version());
  printf("M %lld Np %lld cost_data %.15e cost_reg %.15e x1_norm %.12e x2_norm %.12e\n", (long long)M, (long long)Np,
         cd, cr, sqrt(s1), sqrt(s2));
  emba_lm_settings_t s;
  s.max_num_iter = 50; s.tol_fun = 1e-3; s.num_times_tol_fun_sat = 2; s.use_cg = 0; s.cost_type = EMBA_COST_QUADRATIC;
  s.eta = 1.0; s.thres_valid_pixel = 5; s.damping_factor = 1.0; s.alpha = 5.0; s.first_time_window = 1;
  emba_lm_log_t log[64];
  int32_t nlog = 0;
  double fcost = 0;
  CHECK(emba_solve_time_window(h, &s, log, 64, &nlog, &fcost));
  printf("lm_solves %d final_cost %.15e accepts", nlog, fcost);
  for (int i = 0; i < nlog && i < 64; i++) printf(" %d", log[i].accepted);
  printf("\n");
  double* q = (double*)malloc(sizeof(double) * 4 * n);
  CHECK(emba_get_state(h, EMBA_STATE_CURRENT, q, NULL, NULL));
  printf("q_last %.15e %.15e %.15e %.15e\n", q[4 * (n - 1)], q[4 * (n - 1) + 1], q[4 * (n - 1) + 2], q[4 * (n - 1) + 3]);
  emba_destroy(h);
  free(lut); free(x1); free(x2); free(q); free(ev); free(st);
  return 0;
}
